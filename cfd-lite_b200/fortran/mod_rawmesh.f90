!
!  mod_rawmesh.f90 -- reads the raw mesh file written by cfdl_rawmesh_write (include/cfdl.h) and fills
!  the members of geometry_t / mg_lvl that src/setup/cell_input.f90:36-99 fills from the CGNS library
!  (cgns_db_open, cg_zone_read_f, cg_section_read_f, cgns_element_info, cgns_cell_data of
!  src/modules/mod_cgns.f90:16-148).  With it the reference's driver builds and runs without
!  cgnslib 3.2.1 / HDF5: in cell_input.f90 replace the block from `cg%file=cgfilename` (:36) to
!  `call cgns_cell_data(...)` (:93) by
!
!      call rawmesh_read(cgfilename, mg_lvl%cellname, nvx, nsec, mg_lvl%esec, mg_lvl%etype, mg_lvl%ne2vx, &
!                        mg_lvl%sectionName, mg_lvl%e2vx, ne2vxmax, ne, nf, nbf, vx2e_size, geom%x, geom%y, geom%z)
!
!  and keep the assignments of :72-83 (nelem = ne+nbf, mg_lvl%nsec = nsec, ...).  Everything after
!  (add_meshds, calc_aip_xyzip_uns, calc_vol_cv_centers_uns, ...) is unchanged.
!
!  File layout (little endian, stream access):
!    character(8) 'CFDLRAW1' | int64 nvx | int64 nelem | int32 nsec | int32 ne2vx_max |
!    nsec x ( character(32) name | int32 etype | int32 first | int32 last ) |
!    real64 x(nvx) | real64 y(nvx) | real64 z(nvx) | int32 e2vx(ne2vx_max, nelem)
!
!  NOT compiled in this repository's container (no Fortran compiler there); the C reader of the same
!  format (cfdl_rawmesh_read) is covered by tests/test_rawmesh.py.
!
module mod_rawmesh
  use iso_fortran_env, only: int32, int64, real64
  use mod_util          ! element_nface, element_nvx (src/modules/mod_util.f90:166-169)
  implicit none
contains

  subroutine rawmesh_read(filename, cellname, nvx, nsec, esec, etype, ne2vx, sectionName, e2vx, ne2vxmax, &
                          ne, nf, nbf, vx2e_size, x, y, z)
    character(len=*), intent(in) :: filename
    character(len=32), intent(out) :: cellname
    integer, intent(out) :: nvx, nsec, ne2vxmax, ne, nf, nbf, vx2e_size
    integer, allocatable, intent(out) :: esec(:,:), etype(:), ne2vx(:), e2vx(:)
    character(len=32), allocatable, intent(out) :: sectionName(:)
    real, allocatable, intent(out) :: x(:), y(:), z(:)     ! real(8) under -fdefault-real-8, like geom%x
    character(len=8) :: magic
    integer(int64) :: nvx8, nelem8
    integer(int32) :: nsec4, w4, t4, first4, last4
    integer(int32), allocatable :: e2vx4(:)
    real(real64), allocatable :: buf(:)
    integer :: u, s, ios, nfaces2, cnt

    open(newunit=u, file=trim(filename), access='stream', form='unformatted', status='old', action='read', iostat=ios)
    if (ios /= 0) then
      write(*,*) 'rawmesh_read: cannot open ', trim(filename)
      stop
    endif
    read(u) magic
    if (magic /= 'CFDLRAW1') then
      write(*,*) 'rawmesh_read: not a raw mesh file: ', trim(filename)
      stop
    endif
    read(u) nvx8, nelem8, nsec4, w4
    nvx = int(nvx8); nsec = int(nsec4); ne2vxmax = int(w4)
    cellname = 'cell'
    allocate(esec(2,nsec), etype(nsec), ne2vx(nsec), sectionName(nsec))
    ne = 0; nbf = 0; nfaces2 = 0; vx2e_size = 0
    do s = 1, nsec
      read(u) sectionName(s), t4, first4, last4
      etype(s) = int(t4); esec(1,s) = int(first4); esec(2,s) = int(last4)
      cnt = esec(2,s) - esec(1,s) + 1
      ne2vx(s) = element_nvx(etype(s))                   ! what cg_npe_f returns for the linear element types
      if (etype(s) >= 10 .and. etype(s) <= 20) then       ! 3-D cells (cell_input.f90:61-63)
        ne = ne + cnt
        nfaces2 = nfaces2 + element_nface(etype(s)) * cnt
      else                                                ! 2-D boundary elements (:65)
        nbf = nbf + cnt
      endif
      vx2e_size = vx2e_size + element_nvx(etype(s)) * cnt ! :69
    end do
    nf = (nfaces2 + nbf) / 2                              ! :71
    if (ne + nbf /= int(nelem8)) then
      write(*,*) 'rawmesh_read: section ranges do not add up to the element count'
      stop
    endif
    allocate(buf(nvx), x(nvx), y(nvx), z(nvx))
    read(u) buf; x = buf
    read(u) buf; y = buf
    read(u) buf; z = buf
    deallocate(buf)
    allocate(e2vx4(ne2vxmax*(ne+nbf)), e2vx(ne2vxmax*(ne+nbf)))
    read(u) e2vx4                                         ! column-major (ne2vxmax, nelem): cgns_element_info's e2vx(j,i)
    e2vx = int(e2vx4)
    deallocate(e2vx4)
    close(u)
  end subroutine rawmesh_read

end module mod_rawmesh
