!
!  mod_gpu_bridge.f90 -- ISO_C_BINDING shim between the unmodified CFD-Lite Fortran driver and
!  libcfdl.so (include/cfdl.h).  It is the reference-side binding a maintainer adds; nothing
!  else of the reference changes except the three call sites listed in INTEGRATION.md.
!
!  Conventions are the ones the reference already uses for its only FFI (src/VTK/mod_vtk.f90:3-29,
!  src/main.f90:42-44): arrays are passed as their first element, reals are real(c_double)
!  (the code is built with -fdefault-real-8), integers are integer(c_int).
!
!  NOT compiled in this repository's container (no Fortran compiler there); build it where
!  gfortran exists:  gfortran -cpp -fdefault-real-8 -ffree-line-length-512 -c mod_gpu_bridge.f90
!
module mod_gpu_bridge
  use iso_c_binding
  use mod_cell          ! geometry_t
  use mod_uvwp          ! uvwp_t
  use mod_properties
  implicit none

  type(c_ptr), save :: cfdl_h = c_null_ptr

  ! field selectors, must match the CFDL_F_* enum of cfdl.h
  integer(c_int), parameter :: F_U=0,F_V=1,F_W=2,F_P=3,F_U0=4,F_V0=5,F_W0=6,F_PC=7, &
                               F_GU=8,F_GV=9,F_GW=10,F_GP=11,F_GPC=12,F_MIP=13,F_MIP0=14, &
                               F_BU=15,F_BV=16,F_BW=17,F_D=18,F_DC=19,F_AP=20,F_B=21,F_ANB=22
  integer(c_int), parameter :: BC_WALL=0, BC_LID=1, BC_SYMMETRY=2
  integer(c_int), parameter :: SOLVER_PARITY=0, SOLVER_MCSGS=1, SOLVER_PCG=2

  interface
    function cfdl_last_error() bind(C,name='cfdl_last_error') result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function
    function cfdl_create(h,ne,nf,nbf,ef2nb_idx,ef2nb_nb,ef2nb_fg,s2g,bs,xc,yc,zc,aip,rip,vol,rho,mu, &
                         nbc,bc_esec,bc_kind,bc_uvw,n_subdomains,g2gf_p,g2gf_idx,device) bind(C,name='cfdl_create') result(ierr)
      import :: c_ptr,c_int,c_double
      type(c_ptr) :: h                                   ! cfdl_handle* (by reference)
      integer(c_int), value :: ne,nf,nbf,nbc,n_subdomains,device
      integer(c_int) :: ef2nb_idx(*),ef2nb_nb(*),ef2nb_fg(*),s2g(*),bs(*),bc_esec(*),bc_kind(*),g2gf_p(*),g2gf_idx(*)
      real(c_double) :: xc(*),yc(*),zc(*),aip(*),rip(*),vol(*),rho(*),mu(*),bc_uvw(*)
      integer(c_int) :: ierr
    end function
    function cfdl_destroy(h) bind(C,name='cfdl_destroy') result(ierr)
      import :: c_ptr,c_int
      type(c_ptr), value :: h
      integer(c_int) :: ierr
    end function
    function cfdl_set_option(h,key,val) bind(C,name='cfdl_set_option') result(ierr)
      import :: c_ptr,c_int,c_double,c_char
      type(c_ptr), value :: h
      character(kind=c_char) :: key(*)
      real(c_double), value :: val
      integer(c_int) :: ierr
    end function
    function cfdl_upload_field(h,field,host) bind(C,name='cfdl_upload_field') result(ierr)
      import :: c_ptr,c_int,c_double
      type(c_ptr), value :: h
      integer(c_int), value :: field
      real(c_double) :: host(*)
      integer(c_int) :: ierr
    end function
    function cfdl_download_field(h,field,host) bind(C,name='cfdl_download_field') result(ierr)
      import :: c_ptr,c_int,c_double
      type(c_ptr), value :: h
      integer(c_int), value :: field
      real(c_double) :: host(*)
      integer(c_int) :: ierr
    end function
    function cfdl_update_boundaries(h) bind(C,name='cfdl_update_boundaries') result(ierr)
      import :: c_ptr,c_int
      type(c_ptr), value :: h
      integer(c_int) :: ierr
    end function
    function cfdl_solve_uvwp(h,dt,nit,hist) bind(C,name='cfdl_solve_uvwp') result(ierr)
      import :: c_ptr,c_int,c_double
      type(c_ptr), value :: h
      real(c_double), value :: dt
      integer(c_int), value :: nit
      real(c_double) :: hist(4,4)                        ! hist(:,k) = it,res_i,res_f,res_max of u,v,w,pc
      integer(c_int) :: ierr
    end function
    function cfdl_step_host(h,dt,nit,apply_bcs,local_numbering,n_in,in_fields,in_ptrs,n_out,out_fields,out_ptrs,hist) &
             bind(C,name='cfdl_step_host') result(ierr)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: h
      real(c_double), value :: dt
      integer(c_int), value :: nit, apply_bcs, local_numbering, n_in, n_out
      integer(c_int) :: in_fields(*), out_fields(*)
      type(c_ptr) :: in_ptrs(*), out_ptrs(*)
      real(c_double) :: hist(4,4)
      integer(c_int) :: ierr
    end function
    function cfdl_host_register(ptr,bytes) bind(C,name='cfdl_host_register') result(ierr)
      import :: c_ptr, c_int, c_int64_t
      type(c_ptr), value :: ptr
      integer(c_int64_t), value :: bytes
      integer(c_int) :: ierr
    end function
    function cfdl_update_time(h) bind(C,name='cfdl_update_time') result(ierr)
      import :: c_ptr,c_int
      type(c_ptr), value :: h
      integer(c_int) :: ierr
    end function
  end interface

contains

  subroutine gpu_check(ierr,where)
    integer(c_int) :: ierr
    character(len=*) :: where
    character(kind=c_char), pointer :: msg(:)
    integer :: i
    if(ierr==0) return
    call c_f_pointer(cfdl_last_error(),msg,[512])
    write(*,'(A)',advance='no') 'cfdl error in '//where//': '
    do i=1,512
      if(msg(i)==c_null_char) exit
      write(*,'(A)',advance='no') msg(i)
    end do
    write(*,*)
    stop   ! the reference's own failure convention (mod_eqn_setup.f90:63-66)
  end subroutine

  ! replaces the device-independent part of construct_physics (mod_physics.f90:52-75): call it
  ! right after construct_physics(phys,geom) in main.f90:38
  subroutine gpu_construct(eqn,prop,geom,n_subdomains,solver_mode)
    type(uvwp_t) :: eqn
    type(properties_t) :: prop
    type(geometry_t) :: geom
    integer :: n_subdomains,solver_mode
    integer(c_int), allocatable :: esec(:),kind(:)
    real(c_double), allocatable :: uvw(:)
    integer(c_int) :: dummy(1)
    integer :: i,nbc,Z

    nbc=size(eqn%bcs)
    allocate(esec(2*nbc),kind(nbc),uvw(3*nbc))
    uvw=0.
    do i=1,nbc
      esec(2*i-1:2*i)=eqn%bcs(i)%esec
      ! the callbacks bound in construct_uvwp (mod_uvwp.f90:73-78)
      if(index(trim(eqn%bcs(i)%name),'top')>0) then
        kind(i)=BC_LID; uvw(3*i-2)=1.
      else
        kind(i)=BC_WALL
      end if
    end do
    Z=2*geom%nf-geom%nbf
    dummy=0
    if(n_subdomains>1) then
      call gpu_check(cfdl_create(cfdl_h,geom%ne,geom%nf,geom%nbf,geom%ef2nb_idx(1),geom%ef2nb(1,1),geom%ef2nb(1,2), &
             geom%mg%fine_lvl%s2g(1),geom%mg%fine_lvl%bs(geom%ne+1),geom%xc(1),geom%yc(1),geom%zc(1),geom%aip(1),geom%rip(1), &
             geom%vol(1),prop%rho(1),prop%mu(1),nbc,esec(1),kind(1),uvw(1),n_subdomains, &
             geom%mg%g2gf(1)%p(1),geom%mg%g2gf(1)%idx(1),0),'cfdl_create')
    else
      call gpu_check(cfdl_create(cfdl_h,geom%ne,geom%nf,geom%nbf,geom%ef2nb_idx(1),geom%ef2nb(1,1),geom%ef2nb(1,2), &
             geom%mg%fine_lvl%s2g(1),geom%mg%fine_lvl%bs(geom%ne+1),geom%xc(1),geom%yc(1),geom%zc(1),geom%aip(1),geom%rip(1), &
             geom%vol(1),prop%rho(1),prop%mu(1),nbc,esec(1),kind(1),uvw(1),1,dummy(1),dummy(1),0),'cfdl_create')
    end if
    call gpu_check(cfdl_set_option(cfdl_h,'solver'//c_null_char,real(solver_mode,c_double)),'cfdl_set_option')
    deallocate(esec,kind,uvw)
  end subroutine

  ! drop-in for `call update_boundaries(phys,geom)` (main.f90:52)
  subroutine gpu_update_boundaries()
    call gpu_check(cfdl_update_boundaries(cfdl_h),'cfdl_update_boundaries')
  end subroutine

  ! drop-in for `call solve_uvwp(...)` (main.f90:56); prints the same four lines as the
  ! reference's solvers (mod_solver.f90:6,184,325)
  subroutine gpu_solve_uvwp(dt,nit)
    real :: dt
    integer :: nit
    real(c_double) :: hist(4,4)
    character(len=16) :: names(4)
    integer :: k
    names=[character(len=16) :: 'u','v','w','pc']
    call gpu_check(cfdl_solve_uvwp(cfdl_h,real(dt,c_double),int(nit,c_int),hist),'cfdl_solve_uvwp')
    do k=1,4
      write(*,"(5x,A16,x,i5,x,15x,es9.3e2,3x,es9.3e2,3x,es9.3e2)") names(k),int(hist(1,k)),hist(2,k),hist(3,k),hist(4,k)
    end do
  end subroutine

  ! page-lock the state arrays of uvwp_t once (after gpu_construct) so that gpu_step_host's transfers
  ! overlap the computation; pageable arrays work too, without the overlap
  subroutine gpu_pin_state(eqn)
    type(uvwp_t), target :: eqn
    call pin(eqn%u); call pin(eqn%v); call pin(eqn%w); call pin(eqn%p)
    call pin(eqn%u0); call pin(eqn%v0); call pin(eqn%w0)
    call pin(eqn%gu); call pin(eqn%gv); call pin(eqn%gw); call pin(eqn%gp); call pin(eqn%gpc)
    call pin(eqn%mip); call pin(eqn%mip0)
  contains
    subroutine pin(a)
      real, target :: a(:)
      call gpu_check(cfdl_host_register(c_loc(a(1)), int(size(a),c_int64_t)*8_c_int64_t),'cfdl_host_register')
    end subroutine
  end subroutine

  ! update_boundaries + solve_uvwp with the HOST arrays of uvwp_t staying authoritative (for drivers
  ! that touch the fields between iterations): one call, transfers overlapped with the computation
  ! inside the library (cfdl_step_host).  Equivalent to upload / gpu_update_boundaries /
  ! gpu_solve_uvwp / download of the thirteen state arrays.
  subroutine gpu_step_host(eqn,dt,nit)
    type(uvwp_t), target :: eqn
    real :: dt
    integer :: nit
    real(c_double) :: hist(4,4)
    integer(c_int) :: fin(13), fout(10)
    type(c_ptr) :: pin(13), pout(10)
    fin =[F_U,F_V,F_W,F_P,F_U0,F_V0,F_W0,F_GU,F_GV,F_GW,F_GP,F_MIP,F_MIP0]
    pin =[c_loc(eqn%u(1)),c_loc(eqn%v(1)),c_loc(eqn%w(1)),c_loc(eqn%p(1)),c_loc(eqn%u0(1)),c_loc(eqn%v0(1)),c_loc(eqn%w0(1)), &
          c_loc(eqn%gu(1)),c_loc(eqn%gv(1)),c_loc(eqn%gw(1)),c_loc(eqn%gp(1)),c_loc(eqn%mip(1)),c_loc(eqn%mip0(1))]
    fout=[F_U,F_V,F_W,F_P,F_GU,F_GV,F_GW,F_GP,F_GPC,F_MIP]
    pout=[c_loc(eqn%u(1)),c_loc(eqn%v(1)),c_loc(eqn%w(1)),c_loc(eqn%p(1)),c_loc(eqn%gu(1)),c_loc(eqn%gv(1)),c_loc(eqn%gw(1)), &
          c_loc(eqn%gp(1)),c_loc(eqn%gpc(1)),c_loc(eqn%mip(1))]
    call gpu_check(cfdl_step_host(cfdl_h,real(dt,c_double),int(nit,c_int),1_c_int,0_c_int,13_c_int,fin,pin,10_c_int,fout,pout,hist), &
                   'cfdl_step_host')
  end subroutine

  ! drop-in for `call update_time(phys)` (main.f90:63)
  subroutine gpu_update_time()
    call gpu_check(cfdl_update_time(cfdl_h),'cfdl_update_time')
  end subroutine

  ! before write_vtubin (main.f90:79,89): bring the fields the writer reads back to the host
  subroutine gpu_download(eqn)
    type(uvwp_t) :: eqn
    call gpu_check(cfdl_download_field(cfdl_h,F_U,eqn%u(1)),'download u')
    call gpu_check(cfdl_download_field(cfdl_h,F_V,eqn%v(1)),'download v')
    call gpu_check(cfdl_download_field(cfdl_h,F_W,eqn%w(1)),'download w')
    call gpu_check(cfdl_download_field(cfdl_h,F_P,eqn%p(1)),'download p')
    call gpu_check(cfdl_download_field(cfdl_h,F_GPC,eqn%gpc(1)),'download gpc')
    call gpu_check(cfdl_download_field(cfdl_h,F_MIP,eqn%mip(1)),'download mip')
    call gpu_check(cfdl_download_field(cfdl_h,F_DC,eqn%dc(1)),'download dc')
  end subroutine

  subroutine gpu_destroy()
    call gpu_check(cfdl_destroy(cfdl_h),'cfdl_destroy')
    cfdl_h=c_null_ptr
  end subroutine

end module mod_gpu_bridge
