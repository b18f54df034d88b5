!> Drop-in replacement of module ice_bergs (src/icebergs.F90:1-66) whose hot path runs in
!! libkid_b200.so.  Keeps the public names and argument lists of the reference
!! (icebergs_init I:92-117, icebergs_run I:5074-5096, icebergs_end I:8152, icebergs_stock_pe
!! I:8102, icebergs_incr_mass I:6046, icebergs_save_restart I:8136) so the stand-alone driver
!! (driver/icebergs_driver.F90:339-344, 386-392, 435, 444) and SIS/SIS2 link against it unchanged.
!!
!! Host-side work that stays in Fortran/FMS: namelist -> KidParams, mpp_define_domains ->
!! KidDomain, get_date/yearday, NetCDF restart I/O (columns <-> kid_set_bergs/kid_get_bergs),
!! diag_manager send_data of the fields fetched with kid_get_grid_field, error_mesg.
!!
!! NOT COMPILED IN THIS REPOSITORY'S CI: the image has no Fortran compiler.  The Python mirror
!! icebergs_b200/api.py drives the same C ABI with Fortran-ordered arrays and is what the
!! parity tests exercise.
module ice_bergs
  use, intrinsic :: iso_c_binding
  use fms_mod, only: error_mesg, FATAL, open_namelist_file, check_nml_error, close_file
  use mpp_mod, only: mpp_pe, mpp_npes, mpp_root_pe, mpp_broadcast
  use mpp_domains_mod, only: domain2D, mpp_define_domains, mpp_get_compute_domain, mpp_get_data_domain, &
                             mpp_get_neighbor_pe, NORTH, SOUTH, EAST, WEST, CYCLIC_GLOBAL_DOMAIN
  use mpp_parameter_mod, only: BGRID_NE, CGRID_NE, AGRID
  use time_manager_mod, only: time_type, get_date
  implicit none ; private

  include 'kid_b200_types.inc'     ! generated from include/kid_b200.h (integration/gen_fortran_types.py)

  public :: icebergs, icebergs_init, icebergs_run, icebergs_end, icebergs_stock_pe, icebergs_incr_mass
  public :: icebergs_save_restart

  !> berg slots per rank (0: sized from the restart file, see icebergs_init); read from &icebergs_nml with the rest
  integer(c_int64_t), save :: kid_capacity = 0

  !> The opaque container the callers hold (type(icebergs), pointer :: bergs)
  type :: icebergs
    type(c_ptr) :: h = c_null_ptr          !< kid_t*
    type(KidParams) :: p
    type(KidDomain) :: dom
    type(domain2D) :: domain
    integer :: isc, iec, jsc, jec
  end type icebergs

  interface
    subroutine kid_default_params(p) bind(C, name='kid_default_params')
      import :: KidParams
      type(KidParams), intent(out) :: p
    end subroutine
    function kid_init(h, p, dom, year, yearday, capacity, lon, lat, wet, dx, dy, area, cos_rot, sin_rot, &
                      ocean_depth, fractional_area) bind(C, name='kid_init') result(rc)
      import :: c_ptr, c_int32_t, c_int64_t, c_double, KidParams, KidDomain
      type(c_ptr), intent(out) :: h
      type(KidParams), intent(in) :: p
      type(KidDomain), intent(in) :: dom
      integer(c_int32_t), value :: year, fractional_area
      real(c_double), value :: yearday
      integer(c_int64_t), value :: capacity
      real(c_double), intent(in) :: lon(*), lat(*), wet(*), dx(*), dy(*), area(*), cos_rot(*), sin_rot(*)
      type(c_ptr), value :: ocean_depth          ! may be c_null_ptr
      integer(c_int32_t) :: rc
    end function
    function kid_run(h, year, yearday, calving, uo, vo, ui, vi, tauxa, tauya, ssh, sst, calving_hflx, cn, hi, &
                     stagger, stress_stagger, sss, mass_berg, ustar_berg, area_berg) bind(C, name='kid_run') result(rc)
      import :: c_ptr, c_int32_t, c_double
      type(c_ptr), value :: h
      integer(c_int32_t), value :: year, stagger, stress_stagger
      real(c_double), value :: yearday
      real(c_double), intent(inout) :: calving(*), calving_hflx(*)
      real(c_double), intent(in) :: uo(*), vo(*), ui(*), vi(*), tauxa(*), tauya(*), ssh(*), sst(*), cn(*), hi(*)
      type(c_ptr), value :: sss, mass_berg, ustar_berg, area_berg   ! optional: c_null_ptr when absent
      integer(c_int32_t) :: rc
    end function
    function kid_end(h) bind(C, name='kid_end') result(rc)
      import :: c_ptr, c_int32_t
      type(c_ptr), intent(inout) :: h
      integer(c_int32_t) :: rc
    end function
    function kid_set_bergs(h, n, cols) bind(C, name='kid_set_bergs') result(rc)
      import :: c_ptr, c_int32_t, c_int64_t, KidBergColumns
      type(c_ptr), value :: h
      integer(c_int64_t), value :: n
      type(KidBergColumns), intent(in) :: cols
      integer(c_int32_t) :: rc
    end function
    function kid_get_bergs(h, n, cols, include_halo) bind(C, name='kid_get_bergs') result(rc)
      import :: c_ptr, c_int32_t, c_int64_t, KidBergColumns
      type(c_ptr), value :: h
      integer(c_int64_t), intent(inout) :: n
      type(KidBergColumns), intent(inout) :: cols
      integer(c_int32_t), value :: include_halo
      integer(c_int32_t) :: rc
    end function
    function kid_set_calving_state(h, stored_ice, stored_heat, counter) bind(C, name='kid_set_calving_state') result(rc)
      import :: c_ptr, c_int32_t
      type(c_ptr), value :: h, stored_ice, stored_heat, counter
      integer(c_int32_t) :: rc
    end function
    function kid_set_calving_rmean(h, rmean_calving, rmean_calving_hflx) bind(C, name='kid_set_calving_rmean') result(rc)
      import :: c_ptr, c_int32_t
      type(c_ptr), value :: h, rmean_calving, rmean_calving_hflx      ! c_null_ptr: the variable is not in calving.res.nc (IO:1517-1534)
      integer(c_int32_t) :: rc
    end function
    function kid_get_grid_field(h, field_id, out) bind(C, name='kid_get_grid_field') result(rc)
      import :: c_ptr, c_int32_t, c_double
      type(c_ptr), value :: h
      integer(c_int32_t), value :: field_id
      real(c_double), intent(out) :: out(*)
      integer(c_int32_t) :: rc
    end function
    function kid_stock(h, index, value) bind(C, name='kid_stock') result(rc)
      import :: c_ptr, c_int32_t, c_double
      type(c_ptr), value :: h
      integer(c_int32_t), value :: index
      real(c_double), intent(out) :: value
      integer(c_int32_t) :: rc
    end function
    function kid_incr_mass(h, mass) bind(C, name='kid_incr_mass') result(rc)
      import :: c_ptr, c_int32_t, c_double
      type(c_ptr), value :: h
      real(c_double), intent(inout) :: mass(*)
      integer(c_int32_t) :: rc
    end function
    function kid_nccl_unique_id(buf, nbytes) bind(C, name='kid_nccl_unique_id') result(rc)
      import :: c_char, c_int32_t
      character(kind=c_char), intent(out) :: buf(*)
      integer(c_int32_t), value :: nbytes
      integer(c_int32_t) :: rc
    end function
    function kid_nccl_init(comm, id, nbytes, nranks, rank, device) bind(C, name='kid_nccl_init') result(rc)
      import :: c_ptr, c_char, c_int32_t
      type(c_ptr), intent(out) :: comm
      character(kind=c_char), intent(in) :: id(*)
      integer(c_int32_t), value :: nbytes, nranks, rank, device
      integer(c_int32_t) :: rc
    end function
    function kid_count_bergs(h, n) bind(C, name='kid_count_bergs') result(rc)
      import :: c_ptr, c_int32_t, c_int64_t
      type(c_ptr), value :: h
      integer(c_int64_t), intent(out) :: n
      integer(c_int32_t) :: rc
    end function
    function kid_get_calving_state(h, stored_ice, stored_heat, counter) bind(C, name='kid_get_calving_state') result(rc)
      import :: c_ptr, c_int32_t, c_double
      type(c_ptr), value :: h
      real(c_double), intent(out) :: stored_ice(*), stored_heat(*)
      integer(c_int32_t), intent(out) :: counter(*)
      integer(c_int32_t) :: rc
    end function
    function kid_set_bonds(h, n, cols) bind(C, name='kid_set_bonds') result(rc)
      import :: c_ptr, c_int32_t, c_int64_t, KidBondColumns
      type(c_ptr), value :: h
      integer(c_int64_t), value :: n
      type(KidBondColumns), intent(in) :: cols
      integer(c_int32_t) :: rc
    end function
    function kid_get_bonds(h, n, cols) bind(C, name='kid_get_bonds') result(rc)
      import :: c_ptr, c_int32_t, c_int64_t, KidBondColumns
      type(c_ptr), value :: h
      integer(c_int64_t), intent(inout) :: n
      type(KidBondColumns), intent(inout) :: cols
      integer(c_int32_t) :: rc
    end function
    !> record_posn F:5328: called when the shim's own sample_traj (I:5173-5178) says so
    function kid_record_posn(h) bind(C, name='kid_record_posn') result(rc)
      import :: c_ptr, c_int32_t
      type(c_ptr), value :: h
      integer(c_int32_t) :: rc
    end function
    function kid_get_trajectory(h, n, cols, clear) bind(C, name='kid_get_trajectory') result(rc)
      import :: c_ptr, c_int32_t, c_int64_t, KidTrajColumns
      type(c_ptr), value :: h
      integer(c_int64_t), intent(inout) :: n
      type(KidTrajColumns), intent(inout) :: cols
      integer(c_int32_t), value :: clear
      integer(c_int32_t) :: rc
    end function
    function kid_last_error(h) bind(C, name='kid_last_error') result(msg)
      import :: c_ptr
      type(c_ptr), value :: h
      type(c_ptr) :: msg
    end function
  end interface

contains

  !> error_mesg(..., FATAL) with the library's message (the library never aborts by itself)
  subroutine check(bergs_h, rc, routine)
    type(c_ptr), intent(in) :: bergs_h
    integer(c_int32_t), intent(in) :: rc
    character(len=*), intent(in) :: routine
    character(kind=c_char), pointer :: s(:)
    character(len=512) :: msg
    integer :: k
    if (rc == KID_OK) return
    call c_f_pointer(kid_last_error(bergs_h), s, [512])
    msg = ' '
    do k = 1, 512
      if (s(k) == c_null_char) exit
      msg(k:k) = s(k)
    enddo
    call error_mesg(routine, trim(msg), FATAL)
  end subroutine check

  !> I:92-117.  The namelist read (F:825-879) fills type(KidParams); the FMS domain fills KidDomain.
  subroutine icebergs_init(bergs, gni, gnj, layout, io_layout, axes, dom_x_flags, dom_y_flags, &
                           dt, Time, ice_lon, ice_lat, ice_wet, ice_dx, ice_dy, ice_area, &
                           cos_rot, sin_rot, ocean_depth, maskmap, fractional_area)
    type(icebergs), pointer :: bergs
    integer, intent(in) :: gni, gnj, layout(2), io_layout(2), axes(2), dom_x_flags, dom_y_flags
    real, intent(in) :: dt
    type(time_type), intent(in) :: Time
    real, dimension(:,:), intent(in) :: ice_lon, ice_lat, ice_wet, ice_dx, ice_dy, ice_area, cos_rot, sin_rot
    real, dimension(:,:), intent(in), optional, target :: ocean_depth
    logical, intent(in), optional :: maskmap(:,:), fractional_area
    integer :: iyr, imon, iday, ihr, imin, isec, frac
    integer(c_int64_t) :: capacity
    character(kind=c_char) :: uid(KID_NCCL_UNIQUE_ID_BYTES)
    type(c_ptr) :: depth_ptr

    allocate(bergs)
    call kid_default_params(bergs%p)
    call read_icebergs_nml(bergs%p)           ! namelist icebergs_nml -> KidParams fields of the same name (F:825-856);
                                              ! also reads the one entry the library adds, kid_capacity (module variable, default 0)
    bergs%p%dt = dt
    ! domain exactly as the reference defines it (F:915-930)
    call mpp_define_domains((/1,gni,1,gnj/), layout, bergs%domain, maskmap=maskmap, xflags=dom_x_flags, &
                            yflags=dom_y_flags, xhalo=bergs%p%halo, yhalo=bergs%p%halo, name='diamond')
    call mpp_get_compute_domain(bergs%domain, bergs%dom%isc, bergs%dom%iec, bergs%dom%jsc, bergs%dom%jec)
    call mpp_get_data_domain(bergs%domain, bergs%dom%isd, bergs%dom%ied, bergs%dom%jsd, bergs%dom%jed)
    call mpp_get_neighbor_pe(bergs%domain, NORTH, bergs%dom%pe_N)
    call mpp_get_neighbor_pe(bergs%domain, SOUTH, bergs%dom%pe_S)
    call mpp_get_neighbor_pe(bergs%domain, EAST, bergs%dom%pe_E)
    call mpp_get_neighbor_pe(bergs%domain, WEST, bergs%dom%pe_W)
    bergs%dom%gni = gni ; bergs%dom%gnj = gnj
    bergs%dom%cyclic_x = merge(1, 0, iand(dom_x_flags, CYCLIC_GLOBAL_DOMAIN) /= 0)
    bergs%dom%cyclic_y = merge(1, 0, iand(dom_y_flags, CYCLIC_GLOBAL_DOMAIN) /= 0)
    bergs%dom%fold_north = merge(1, 0, dom_y_flags == FOLD_NORTH_EDGE)      ! the tripolar grid, F:933
    bergs%dom%rank = mpp_pe() - mpp_root_pe() ; bergs%dom%nranks = mpp_npes()
    bergs%dom%layout_x = layout(1) ; bergs%dom%layout_y = layout(2)
    bergs%dom%device = local_device_ordinal()  ! e.g. rank modulo GPUs per node
    bergs%dom%comm_kind = KID_COMM_NCCL
    bergs%dom%nccl_comm = c_null_ptr
    if (mpp_npes() > 1) then                   ! NCCL bootstrap over the MPI the host already has
      if (mpp_pe() == mpp_root_pe()) call check(c_null_ptr, kid_nccl_unique_id(uid, KID_NCCL_UNIQUE_ID_BYTES), 'KID, icebergs_init')
      call mpp_broadcast(uid, KID_NCCL_UNIQUE_ID_BYTES, mpp_root_pe())
      call check(c_null_ptr, kid_nccl_init(bergs%dom%nccl_comm, uid, KID_NCCL_UNIQUE_ID_BYTES, mpp_npes(), &
                                           bergs%dom%rank, bergs%dom%device), 'KID, icebergs_init')
    endif
    call get_date(Time, iyr, imon, iday, ihr, imin, isec)
    frac = 0 ; if (present(fractional_area)) frac = merge(1, 0, fractional_area)
    depth_ptr = c_null_ptr ; if (present(ocean_depth)) depth_ptr = c_loc(ocean_depth)
    ! the berg store is sized once: four times what this rank's restart file holds (the length of its "i" dimension,
    ! IO:640-660), at least 2**20 -- calving, footloose children and arrivals are appended to it (KID_ERR_CAPACITY when it
    ! overflows; the namelist entry kid_capacity overrides)
    capacity = max(int(2, c_int64_t)**20, 4_c_int64_t * restart_berg_count(bergs%domain))
    if (kid_capacity > 0) capacity = kid_capacity
    call check(bergs%h, kid_init(bergs%h, bergs%p, bergs%dom, iyr, yearday(imon, iday, ihr, imin, isec), &
               capacity, ice_lon, ice_lat, ice_wet, ice_dx, ice_dy, ice_area, cos_rot, sin_rot, &
               depth_ptr, frac), 'KID, icebergs_init')
    ! read_restart_calving / read_restart_bergs (IO:606-975, IO:1432-1530) stay on the host: the NetCDF
    ! columns go to kid_set_calving_state / kid_set_bergs unchanged (names as IO:261-337)
    call read_restart_into_library(bergs, Time)
  end subroutine icebergs_init

  !> I:5074-5096
  subroutine icebergs_run(bergs, time, calving, uo, vo, ui, vi, tauxa, tauya, ssh, sst, calving_hflx, cn, hi, &
                          stagger, stress_stagger, sss, mass_berg, ustar_berg, area_berg)
    type(icebergs), pointer :: bergs
    type(time_type), intent(in) :: time
    real, dimension(:,:), intent(inout) :: calving, calving_hflx
    real, dimension(:,:), intent(in) :: uo, vo, ui, vi, tauxa, tauya, ssh, sst, cn, hi
    integer, optional, intent(in) :: stagger, stress_stagger
    real, dimension(:,:), optional, intent(in), target :: sss
    real, dimension(:,:), optional, pointer :: mass_berg, ustar_berg, area_berg
    integer :: iyr, imon, iday, ihr, imin, isec, vel_stagger, str_stagger
    type(c_ptr) :: p_sss, p_m, p_u, p_a
    vel_stagger = BGRID_NE ; if (present(stagger)) vel_stagger = stagger
    str_stagger = vel_stagger ; if (present(stress_stagger)) str_stagger = stress_stagger
    p_sss = c_null_ptr ; if (present(sss)) p_sss = c_loc(sss)
    p_m = c_null_ptr ; p_u = c_null_ptr ; p_a = c_null_ptr
    if (present(mass_berg)) then ; if (associated(mass_berg)) p_m = c_loc(mass_berg) ; endif
    if (present(ustar_berg)) then ; if (associated(ustar_berg)) p_u = c_loc(ustar_berg) ; endif
    if (present(area_berg)) then ; if (associated(area_berg)) p_a = c_loc(area_berg) ; endif
    call get_date(time, iyr, imon, iday, ihr, imin, isec)
    call check(bergs%h, kid_run(bergs%h, iyr, yearday(imon, iday, ihr, imin, isec), calving, uo, vo, ui, vi, &
               tauxa, tauya, ssh, sst, calving_hflx, cn, hi, kid_stagger(vel_stagger), kid_stagger(str_stagger), &
               p_sss, p_m, p_u, p_a), 'KID, icebergs_run')
    ! diagnostics: kid_get_grid_field(KID_FLD_*) -> send_data, as I:5520-5650
  end subroutine icebergs_run

  integer(c_int32_t) function kid_stagger(s)
    integer, intent(in) :: s
    kid_stagger = KID_BGRID_NE
    if (s == CGRID_NE) kid_stagger = KID_CGRID_NE
    if (s == AGRID) kid_stagger = KID_AGRID
  end function kid_stagger

  !> F:4431-4441
  real function yearday(imon, iday, ihr, imin, isec)
    integer, intent(in) :: imon, iday, ihr, imin, isec
    yearday = float(imon-1)*31. + float(iday-1) + (float(ihr) + (float(imin) + float(isec)/60.)/60.)/24.
  end function yearday

  !> I:8152: write restarts from kid_get_bergs columns (IO:261-337), then free the device state
  subroutine icebergs_end(bergs)
    type(icebergs), pointer :: bergs
    if (.not. associated(bergs)) return
    call icebergs_save_restart(bergs)
    call check(bergs%h, kid_end(bergs%h), 'KID, icebergs_end')
    deallocate(bergs)
  end subroutine icebergs_end

  !> I:8102
  subroutine icebergs_stock_pe(bergs, index, value)
    type(icebergs), pointer :: bergs
    integer, intent(in) :: index
    real, intent(out) :: value
    call check(bergs%h, kid_stock(bergs%h, int(index - 1, c_int32_t), value), 'KID, icebergs_stock_pe')
  end subroutine icebergs_stock_pe

  !> I:6046
  subroutine icebergs_incr_mass(bergs, mass, Time)
    type(icebergs), pointer :: bergs
    real, dimension(bergs%isc:bergs%iec, bergs%jsc:bergs%jec), intent(inout) :: mass
    type(time_type), intent(in), optional :: Time
    call check(bergs%h, kid_incr_mass(bergs%h, mass), 'KID, icebergs_incr_mass')
  end subroutine icebergs_incr_mass

  !> &icebergs_nml (F:825-856) -> KidParams.  Declarations, the namelist statement and the copies are generated from
  !! include/kid_b200.h (integration/kid_b200_nml.inc), so every field the library knows is a namelist entry of the
  !! same name; entries of the reference's namelist that never reach the hot path (verbose, budget, restart and
  !! trajectory file names, ...) are read by the FMS-side code that still owns them.
  subroutine read_icebergs_nml(p)
    type(KidParams), intent(inout) :: p
    integer :: iunit, ierr
#define KID_NML_COPY_IN
#include "kid_b200_nml.inc"
#undef KID_NML_COPY_IN
    iunit = open_namelist_file()
    read(iunit, icebergs_nml, iostat=ierr)
    ierr = check_nml_error(ierr, 'icebergs_nml')
    call close_file(iunit)
    call copy_out()
  contains
    subroutine copy_out()
#define KID_NML_COPY_OUT
#define KID_NML_NO_DECLS
#include "kid_b200_nml.inc"
#undef KID_NML_COPY_OUT
    end subroutine copy_out
  end subroutine read_icebergs_nml

  !> rank -> CUDA device of this node: ranks are placed round-robin over the GPUs (KID_GPUS_PER_NODE, default 8)
  function local_device_ordinal() result(dev)
    integer(c_int32_t) :: dev
    character(len=16) :: env
    integer :: ngpu, stat
    ngpu = 8
    call get_environment_variable('KID_GPUS_PER_NODE', env, status=stat)
    if (stat == 0) read(env, *, iostat=stat) ngpu
    if (ngpu < 1) ngpu = 1
    dev = int(mod(mpp_pe() - mpp_root_pe(), ngpu), c_int32_t)
  end function local_device_ordinal

  !> read_restart_calving / read_restart_bergs / read_restart_bonds (IO:606-975, IO:1282-1530): the NetCDF / FMS I/O
  !! stays on the host exactly as in icebergs_fmsio.F90; instead of create_iceberg + add_new_berg_to_list per record the
  !! columns it reads go to the library in one call each.
  subroutine read_restart_into_library(bergs, Time)
    type(icebergs), pointer :: bergs
    type(time_type), intent(in) :: Time
    type(KidBergColumns) :: c
    type(KidBondColumns) :: cb
    integer(c_int64_t) :: nb, nbonds
    real, allocatable, target :: stored_ice(:,:,:), stored_heat(:,:)
    integer(c_int32_t), allocatable, target :: counter(:,:)
    ! calving.res.nc: FMS hands back the compute domain; the library holds the data domain (halos zero until the first
    ! halo update of icebergs_run, I:5203)
    allocate(stored_ice(bergs%dom%isd:bergs%dom%ied, bergs%dom%jsd:bergs%dom%jed, KID_NCLASSES), &
             stored_heat(bergs%dom%isd:bergs%dom%ied, bergs%dom%jsd:bergs%dom%jed), &
             counter(bergs%dom%isd:bergs%dom%ied, bergs%dom%jsd:bergs%dom%jed))
    stored_ice = 0. ; stored_heat = 0. ; counter = 0
    call read_restart_calving_fields(bergs%domain, stored_ice, stored_heat, counter)        ! IO:1432-1530, unchanged
    call check(bergs%h, kid_set_calving_state(bergs%h, stored_ice, stored_heat, counter), 'KID, read_restart_calving')
    ! rmean_calving / rmean_calving_hflx (IO:1517-1534): handed over only when the file holds them (a mean that is not
    ! set starts from the first calving field icebergs_run sees, I:6010-6017)
    if (bergs%p%tau_calving > 0.) call read_restart_calving_rmean_into_library(bergs)
    ! icebergs.res.nc: one allocatable array per variable (IO:606-760 reads them exactly like this), handed over by address
    call read_restart_berg_columns(bergs%domain, nb, c)                                        ! IO:606-975 without the list insertion
    call check(bergs%h, kid_set_bergs(bergs%h, nb, c), 'KID, read_restart_bergs')
    if (bergs%p%iceberg_bonds_on /= 0) then
      call read_restart_bond_columns(bergs%domain, nbonds, cb)                                 ! IO:1282-1430; nbonds = 0 without a file
      call check(bergs%h, kid_set_bonds(bergs%h, nbonds, cb), 'KID, read_restart_bonds')       ! (initialize_iceberg_bonds I:356 when empty)
    endif
  end subroutine read_restart_into_library

  !> rmean_calving, rmean_calving_hflx of calving.res.nc (IO:568-569 written, fms2io:1517-1534 read) -> kid_set_calving_rmean
  subroutine read_restart_calving_rmean_into_library(bergs)
    type(icebergs), pointer :: bergs
    real, allocatable, target :: rmean(:,:), rmean_hflx(:,:)
    logical :: has_rmean, has_rmean_hflx
    type(c_ptr) :: pa, pb
    allocate(rmean(bergs%dom%isd:bergs%dom%ied, bergs%dom%jsd:bergs%dom%jed), &
             rmean_hflx(bergs%dom%isd:bergs%dom%ied, bergs%dom%jsd:bergs%dom%jed))
    rmean = 0. ; rmean_hflx = 0.
    call read_restart_calving_rmean_fields(bergs%domain, rmean, has_rmean, rmean_hflx, has_rmean_hflx)   ! variable_exists + read_data
    pa = c_null_ptr ; pb = c_null_ptr
    if (has_rmean) pa = c_loc(rmean)
    if (has_rmean_hflx) pb = c_loc(rmean_hflx)
    call check(bergs%h, kid_set_calving_rmean(bergs%h, pa, pb), 'KID, read_restart_calving')
  end subroutine read_restart_calving_rmean_into_library

  !> I:8136: write_restart_bergs / write_restart_bonds / write_restart_calving (IO:261-566) from the library's columns
  subroutine icebergs_save_restart(bergs, time_stamp)
    type(icebergs), pointer :: bergs
    character(len=*), intent(in), optional :: time_stamp
    type(KidBergColumns) :: c
    type(KidBondColumns) :: cb
    integer(c_int64_t) :: nb, nbonds
    real, allocatable, target :: stored_ice(:,:,:), stored_heat(:,:)
    integer(c_int32_t), allocatable, target :: counter(:,:)
    if (.not. associated(bergs)) return
    call check(bergs%h, kid_count_bergs(bergs%h, nb), 'KID, icebergs_save_restart')
    call allocate_berg_columns(c, nb)                     ! one array per variable of icebergs.res.nc (IO:261-337)
    call check(bergs%h, kid_get_bergs(bergs%h, nb, c, 0_c_int32_t), 'KID, icebergs_save_restart')
    call write_restart_berg_columns(bergs%domain, nb, c, time_stamp)          ! IO:261-430 from arrays instead of the list walk
    if (bergs%p%iceberg_bonds_on /= 0) then
      nbonds = 0
      call check(bergs%h, kid_get_bonds(bergs%h, nbonds, cb), 'KID, icebergs_save_restart')   ! count only (NULL columns)
      call allocate_bond_columns(cb, nbonds)
      call check(bergs%h, kid_get_bonds(bergs%h, nbonds, cb), 'KID, icebergs_save_restart')
      call write_restart_bond_columns(bergs%domain, nbonds, cb, time_stamp)   ! IO:432-560
    endif
    allocate(stored_ice(bergs%dom%isd:bergs%dom%ied, bergs%dom%jsd:bergs%dom%jed, KID_NCLASSES), &
             stored_heat(bergs%dom%isd:bergs%dom%ied, bergs%dom%jsd:bergs%dom%jed), &
             counter(bergs%dom%isd:bergs%dom%ied, bergs%dom%jsd:bergs%dom%jed))
    call check(bergs%h, kid_get_calving_state(bergs%h, stored_ice, stored_heat, counter), 'KID, icebergs_save_restart')
    call write_restart_calving_fields(bergs%domain, stored_ice, stored_heat, counter, time_stamp)   ! IO:564-566
    ! (with tau_calving > 0 the same file also carries kid_get_calving_rmean's two arrays, IO:568-569)
  end subroutine icebergs_save_restart

  ! restart_berg_count(domain): the length of the unlimited dimension "i" of this rank's icebergs.res.nc (what
  ! read_restart_bergs IO:640-660 obtains before it loops over the records; 0 without a file).
  ! read_restart_calving_fields / read_restart_berg_columns / read_restart_bond_columns and their write_* counterparts,
  ! allocate_berg_columns / allocate_bond_columns: the bodies of the routines of icebergs_fmsio.F90 cited above with the
  ! per-record create_iceberg / list walk replaced by the array they already read into or write from (the Python
  ! equivalents, tested against the same files, are icebergs_b200/restart_io.py).
end module ice_bergs
