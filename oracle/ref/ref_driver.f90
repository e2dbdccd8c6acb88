!
!  ref_driver.f90 -- TEST INFRASTRUCTURE: drives the UNMODIFIED reference modules (linked from /root/reference/src by
!  oracle/build_ref.sh) through the loop of src/main.f90:50-63 and dumps what the parity tests compare, in binary
!  (the reference prints the residual history with 4 digits only, src/modules/mod_solver.f90:6):
!
!     ref_dump.bin : int32 ne, nf, nbf, nsolves | real64 hist(4,nsolves) = (it, res_i, res_f, res_max) per linear solve
!                    | real64 u(ne+nbf), v, w, p | real64 mip(nf)
!
!  The history records come from a one-line hook that build_ref.sh appends after each `write(*,oformat)` of
!  mod_solver.f90 in its scratch copy (`call ref_hist_push(it,res_i,res_f,res_max)`); nothing else of the
!  reference is touched.  VTK / Catalyst output is not built.
!
!  usage: ref_driver <mesh.raw> <ntstep> <ncoef> <n_subdomains>
!
module ref_hist
  implicit none
  integer :: nrec = 0
  real, allocatable :: rec(:,:)
contains
  subroutine ref_hist_push(it, res_i, res_f, res_max)
    integer :: it
    real :: res_i, res_f, res_max
    real, allocatable :: tmp(:,:)
    if (.not. allocated(rec)) allocate(rec(4,1024))
    if (nrec == size(rec,2)) then
      allocate(tmp(4,2*nrec)); tmp(:,1:nrec) = rec; call move_alloc(tmp, rec)
    end if
    nrec = nrec + 1
    rec(:,nrec) = [real(it), res_i, res_f, res_max]
  end subroutine
end module

program ref_driver
  use mod_cell
  use mod_physics
  use mod_solver
  use ref_hist
  implicit none
  character(len=180) :: filename
  character(len=32) :: arg
  type(geometry_t) :: geom
  type(phys_t) :: phys
  integer :: tstep, icoef, u

  call get_command_argument(1, filename)
  call get_command_argument(2, arg); read(arg,*) phys%ntstep
  call get_command_argument(3, arg); read(arg,*) phys%ncoef
  call get_command_argument(4, arg); read(arg,*) phys%n_subdomains
  call cell_input(geom, filename, phys%n_subdomains)   ! build_ref.sh's scratch copy reads the raw mesh file (mod_rawmesh)
  call construct_physics(phys, geom)
  do tstep = 1, phys%ntstep                             ! src/main.f90:50-63
    do icoef = 1, phys%ncoef
      call update_boundaries(phys, geom)
      call solve_uvwp(phys%uvwp, phys%prop, geom, phys%dt, phys%nit, phys%ap, phys%anb, phys%b, phys%phic, &
                      phys%subdomain, phys%intf, phys%n_subdomains)
    end do
    call update_time(phys)
  end do
  open(newunit=u, file='ref_dump.bin', access='stream', form='unformatted', status='replace')
  write(u) int(geom%ne,4), int(geom%nf,4), int(geom%nbf,4), int(nrec,4)
  write(u) rec(:,1:nrec)
  write(u) phys%uvwp%u, phys%uvwp%v, phys%uvwp%w, phys%uvwp%p, phys%uvwp%mip
  close(u)
end program
