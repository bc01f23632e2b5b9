!> ISO_C_BINDING declarations of the DEVICE-RESIDENT entry points of liblesgo_cuda.so
!> (include/lesgo_gpu.h): whole-step stepping, the Lagrangian scale-dependent model switches,
!> actuator disks, running time averages and the restart file.  Companion of lesgo_gpu_mod.f90
!> (same context `gpu_ctx`, same error path `gpu_check`); the table in INTEGRATION.md says which
!> reference call site each one replaces.
!> (Not compiled in the build container: it has no Fortran compiler.  The derived types below
!> mirror the C structs member by member; tests/test_abi.py checks the ctypes mirror of the same
!> structs against the header.)
module lesgo_gpu_resident_mod
use iso_c_binding
use lesgo_gpu_mod, only : gpu_ctx, gpu_check
implicit none
save
public

!> enum lesgo_gpu_field
integer(c_int), parameter :: LG_U = 0, LG_V = 1, LG_W = 2, LG_DUDX = 3, LG_DUDY = 4, LG_DUDZ = 5,           &
    LG_DVDX = 6, LG_DVDY = 7, LG_DVDZ = 8, LG_DWDX = 9, LG_DWDY = 10, LG_DWDZ = 11, LG_RHSX = 12,            &
    LG_RHSY = 13, LG_RHSZ = 14, LG_RHSX_F = 15, LG_RHSY_F = 16, LG_RHSZ_F = 17, LG_P = 18, LG_DPDX = 19,     &
    LG_DPDY = 20, LG_DPDZ = 21, LG_DIVTX = 22, LG_DIVTY = 23, LG_DIVTZ = 24, LG_TXX = 25, LG_TXY = 26,       &
    LG_TXZ = 27, LG_TYY = 28, LG_TYZ = 29, LG_TZZ = 30, LG_F_LM = 31, LG_F_MM = 32, LG_F_QN = 33,            &
    LG_F_NN = 34, LG_CS_OPT2 = 35, LG_FXA = 36, LG_FYA = 37, LG_FZA = 38

!> struct lesgo_gpu_step_params
type, bind(c) :: lesgo_gpu_step_params
    real(c_double) :: dt, tadv1, tadv2
    real(c_double) :: mean_p_force_x, mean_p_force_y
    real(c_double) :: ubot, utop, nu_molec_nd
    integer(c_int) :: first_step, mode
    integer(c_int) :: sgs_model, ifilter
    real(c_double) :: Co, wall_damp_exp, vonk, zo
    integer(c_int) :: lasd_cs_init, lasd_update, lasd_init_F
    real(c_double) :: lagran_dt
    integer(c_int) :: turbines
    real(c_double) :: turbines_eps
end type lesgo_gpu_step_params

!> struct lesgo_gpu_turbine: nodes -> c_loc(wind_farm%turbine(s)%nodes_c), an integer(c_int) array (3, num_nodes)
!> holding transpose(%nodes(1:num_nodes, 1:3)); ind -> c_loc(%ind)
type, bind(c) :: lesgo_gpu_turbine
    integer(c_int) :: num_nodes
    type(c_ptr) :: nodes
    type(c_ptr) :: ind
    real(c_double) :: nhat(3)
    real(c_double) :: Ct_prime, dia, M, u_d_T
end type lesgo_gpu_turbine

interface
    integer(c_int) function lesgo_gpu_upload(ctx, field, host) bind(c, name='lesgo_gpu_upload')
        import
        type(c_ptr), value :: ctx
        integer(c_int), value :: field
        real(c_double), intent(in) :: host(*)
    end function
    integer(c_int) function lesgo_gpu_download(ctx, field, host) bind(c, name='lesgo_gpu_download')
        import
        type(c_ptr), value :: ctx
        integer(c_int), value :: field
        real(c_double), intent(out) :: host(*)
    end function
    integer(c_int) function lesgo_gpu_step(ctx, sp) bind(c, name='lesgo_gpu_step')
        import
        type(c_ptr), value :: ctx
        type(lesgo_gpu_step_params), intent(in) :: sp
    end function
    integer(c_int) function lesgo_gpu_max_cfl(ctx, dt, cfl) bind(c, name='lesgo_gpu_max_cfl')
        import
        type(c_ptr), value :: ctx
        real(c_double), value :: dt
        real(c_double), intent(out) :: cfl
    end function
    integer(c_int) function lesgo_gpu_cfl_dt(ctx, cfl, dt) bind(c, name='lesgo_gpu_cfl_dt')
        import
        type(c_ptr), value :: ctx
        real(c_double), value :: cfl
        real(c_double), intent(out) :: dt
    end function
    integer(c_int) function lesgo_gpu_rmsdiv(ctx, rms) bind(c, name='lesgo_gpu_rmsdiv')
        import
        type(c_ptr), value :: ctx
        real(c_double), intent(out) :: rms
    end function
    integer(c_int) function lesgo_gpu_turbines_init(ctx, nloc, turbines, adm_correction)                    &
        bind(c, name='lesgo_gpu_turbines_init')
        import
        type(c_ptr), value :: ctx
        integer(c_int), value :: nloc, adm_correction
        type(lesgo_gpu_turbine), intent(in) :: turbines(*)
    end function
    !> ind_t(s) -> c_loc(%ind_t), e_theta(s) -> c_loc of a real(c_double) array (3, num_nodes) holding
    !> transpose(%e_theta(1:num_nodes, 1:3))
    integer(c_int) function lesgo_gpu_turbines_rotation(ctx, nloc, ind_t, e_theta, tip_speed_ratio)         &
        bind(c, name='lesgo_gpu_turbines_rotation')
        import
        type(c_ptr), value :: ctx
        integer(c_int), value :: nloc
        type(c_ptr), intent(in) :: ind_t(*), e_theta(*)
        real(c_double), value :: tip_speed_ratio
    end function
    integer(c_int) function lesgo_gpu_turbines_forcing(ctx, eps, u_d, u_d_T, f_n)                           &
        bind(c, name='lesgo_gpu_turbines_forcing')
        import
        type(c_ptr), value :: ctx
        real(c_double), value :: eps
        real(c_double), intent(out) :: u_d(*), u_d_T(*), f_n(*)
    end function
    integer(c_int) function lesgo_gpu_tavg_compute(ctx, dt) bind(c, name='lesgo_gpu_tavg_compute')
        import
        type(c_ptr), value :: ctx
        real(c_double), value :: dt
    end function
    integer(c_int) function lesgo_gpu_tavg_download(ctx, which, host, total_time)                           &
        bind(c, name='lesgo_gpu_tavg_download')
        import
        type(c_ptr), value :: ctx
        integer(c_int), value :: which
        real(c_double), intent(out) :: host(*)
        real(c_double), intent(out) :: total_time
    end function
    integer(c_int) function lesgo_gpu_tavg_reset(ctx) bind(c, name='lesgo_gpu_tavg_reset')
        import
        type(c_ptr), value :: ctx
    end function
    integer(c_int) function lesgo_gpu_checkpoint_write(ctx, fname) bind(c, name='lesgo_gpu_checkpoint_write')
        import
        type(c_ptr), value :: ctx
        character(kind=c_char), intent(in) :: fname(*)
    end function
    integer(c_int) function lesgo_gpu_checkpoint_read(ctx, fname) bind(c, name='lesgo_gpu_checkpoint_read')
        import
        type(c_ptr), value :: ctx
        character(kind=c_char), intent(in) :: fname(*)
    end function
end interface

contains

!> The `if` chain of sgs_stag_util.f90:183-216 and lagrange_Sdep.f90:270,320 as switches of the step:
!> call once per time step before lesgo_gpu_step when sgs_model == 5.
subroutine gpu_lasd_switches(sp, lasd_initialised)
use param, only : jt, jt_total, DYN_init, cs_count, inilag, initu, dt, use_cfl_dt
type(lesgo_gpu_step_params), intent(inout) :: sp
logical, intent(inout) :: lasd_initialised      !< F_LM_MM_init / F_QN_NN_init of lagrange_Sdep.f90:70-71
real(c_double), save :: lagran_dt_acc = 0._c_double
sp%lasd_cs_init = 0; sp%lasd_update = 0; sp%lasd_init_F = 0
! sgs_stag_util.f90:73-82: with use_cfl_dt the Lagrangian time interval is ACCUMULATED, lagran_dt += dt on every
! step from DYN_init - cs_count + 1 on (or initu), and reset after lagrange_Sdep (lagrange_Sdep.f90:430); with a
! fixed time step it is cs_count * dt
if (use_cfl_dt) then
    if (jt >= DYN_init - cs_count + 1 .or. initu) lagran_dt_acc = lagran_dt_acc + dt
    sp%lagran_dt = lagran_dt_acc
else
    sp%lagran_dt = cs_count * dt
end if
if (jt == 1 .and. inilag) then
    sp%lasd_cs_init = 1
else if ((jt >= DYN_init .or. initu) .and. mod(jt_total, cs_count) == 0) then
    sp%lasd_update = 1
    lagran_dt_acc = 0._c_double                  ! lagrange_Sdep.f90:430 (after this step has used sp%lagran_dt)
    if (inilag .and. .not. lasd_initialised .and. (jt == cs_count .or. jt == DYN_init)) then
        sp%lasd_init_F = 1
        lasd_initialised = .true.
    end if
end if
end subroutine gpu_lasd_switches

!> Hand wind_farm over after turbines_init / turbines_nodes (turbines.f90:129-462), and again whenever
!> turbines_nodes re-meshes the disks (dyn_theta1 / dyn_theta2, turbines.f90:506-515).  With use_rotation
!> (turbines.f90:76) the tangential weights and unit vectors go along (turbines.f90:419-429, :456, :607-615).
subroutine gpu_turbines_set(wind_farm, nloc, adm_correction, use_rotation, tip_speed_ratio)
use stat_defs, only : wind_farm_t
type(wind_farm_t), intent(in), target :: wind_farm
integer, intent(in) :: nloc
logical, intent(in) :: adm_correction, use_rotation
real(c_double), intent(in) :: tip_speed_ratio
type(lesgo_gpu_turbine), allocatable :: t(:)
type(c_ptr), allocatable :: pt(:), pe(:)
integer(c_int), allocatable, target :: nodes_c(:,:,:)
real(c_double), allocatable, target :: eth_c(:,:,:)
integer :: s, n, nmax
nmax = 1
do s = 1, nloc
    nmax = max(nmax, wind_farm%turbine(s)%num_nodes)
end do
allocate(t(nloc), pt(nloc), pe(nloc), nodes_c(3, nmax, nloc), eth_c(3, nmax, nloc))
do s = 1, nloc
    n = wind_farm%turbine(s)%num_nodes
    nodes_c(:, 1:n, s) = transpose(wind_farm%turbine(s)%nodes(1:n, 1:3))
    t(s)%num_nodes = n
    t(s)%nodes = c_loc(nodes_c(1, 1, s))
    t(s)%ind = c_loc(wind_farm%turbine(s)%ind(1))
    t(s)%nhat = wind_farm%turbine(s)%nhat
    t(s)%Ct_prime = wind_farm%turbine(s)%Ct_prime
    t(s)%dia = wind_farm%turbine(s)%dia
    t(s)%M = wind_farm%turbine(s)%turb_ind_func%M
    t(s)%u_d_T = wind_farm%turbine(s)%u_d_T
    if (use_rotation) then
        eth_c(:, 1:n, s) = transpose(wind_farm%turbine(s)%e_theta(1:n, 1:3))
        pt(s) = c_loc(wind_farm%turbine(s)%ind_t(1))
        pe(s) = c_loc(eth_c(1, 1, s))
    end if
end do
call gpu_check(lesgo_gpu_turbines_init(gpu_ctx, int(nloc, c_int), t, merge(1_c_int, 0_c_int, adm_correction)),        &
    'lesgo_gpu_turbines_init')
if (use_rotation) call gpu_check(lesgo_gpu_turbines_rotation(gpu_ctx, int(nloc, c_int), pt, pe, tip_speed_ratio),     &
    'lesgo_gpu_turbines_rotation')
end subroutine gpu_turbines_set

!> checkpoint (io.f90:1199-1211) / ic_file (initial.f90:226-239) with the reference's file name
subroutine gpu_checkpoint(fname, writing)
character(*), intent(in) :: fname
logical, intent(in) :: writing
if (writing) then
    call gpu_check(lesgo_gpu_checkpoint_write(gpu_ctx, trim(fname)//c_null_char), 'lesgo_gpu_checkpoint_write')
else
    call gpu_check(lesgo_gpu_checkpoint_read(gpu_ctx, trim(fname)//c_null_char), 'lesgo_gpu_checkpoint_read')
end if
end subroutine gpu_checkpoint

end module lesgo_gpu_resident_mod
