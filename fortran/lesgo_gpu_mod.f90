!> ISO_C_BINDING declarations of liblesgo_cuda.so (include/lesgo_gpu.h) and the one
!> context every replacement routine shares.  This file and its siblings are the
!> reference-side binding: drop them into the LESGO source directory IN PLACE OF
!> derivatives.f90, convec.f90, press_stag_array.f90, tridag_array.f90 and fft.f90
!> (CMakeLists.txt:171-181 lists those as ordinary sources) and link liblesgo_cuda.so.
!> Nothing else in LESGO changes: module names, procedure names and argument lists are
!> those of the reference (derivatives.f90:32, fft.f90:31-37, tridag_array.f90:22-32).
!> (Not compiled in the build container: it has no Fortran compiler.  See INTEGRATION.md.)
module lesgo_gpu_mod
use iso_c_binding
implicit none
save
public

type, bind(c) :: lesgo_gpu_dims
    integer(c_int) :: nx, ny, nz, nz_tot, nproc, coord
    real(c_double) :: L_x, L_y, dz
    integer(c_int) :: lbc_mom, ubc_mom, sgs, device
end type lesgo_gpu_dims

type(c_ptr) :: gpu_ctx = c_null_ptr

interface
    integer(c_int) function lesgo_gpu_create(dims, ctx) bind(c, name='lesgo_gpu_create')
        import :: c_int, c_ptr, lesgo_gpu_dims
        type(lesgo_gpu_dims), intent(in) :: dims
        type(c_ptr), intent(out) :: ctx
    end function
    integer(c_int) function lesgo_gpu_destroy(ctx) bind(c, name='lesgo_gpu_destroy')
        import :: c_int, c_ptr
        type(c_ptr), value :: ctx
    end function
    type(c_ptr) function lesgo_gpu_last_error(ctx) bind(c, name='lesgo_gpu_last_error')
        import :: c_ptr
        type(c_ptr), value :: ctx
    end function
    integer(c_int) function lesgo_gpu_comm_unique_id(id) bind(c, name='lesgo_gpu_comm_unique_id')
        import :: c_int, c_char
        character(kind=c_char), intent(out) :: id(128)
    end function
    integer(c_int) function lesgo_gpu_comm_init(ctx, id) bind(c, name='lesgo_gpu_comm_init')
        import :: c_int, c_ptr, c_char
        type(c_ptr), value :: ctx
        character(kind=c_char), intent(in) :: id(128)
    end function
    integer(c_int) function lesgo_gpu_comm_p2p_export(ctx, blob) bind(c, name='lesgo_gpu_comm_p2p_export')
        import
        type(c_ptr), value :: ctx
        character(kind=c_char) :: blob(128)
    end function
    integer(c_int) function lesgo_gpu_comm_p2p_import(ctx, blobs) bind(c, name='lesgo_gpu_comm_p2p_import')
        import
        type(c_ptr), value :: ctx
        character(kind=c_char), intent(in) :: blobs(*)
    end function
    integer(c_int) function lesgo_gpu_host_register(ctx, host, bytes) bind(c, name='lesgo_gpu_host_register')
        import
        type(c_ptr), value :: ctx
        type(c_ptr), value :: host
        integer(c_size_t), value :: bytes
    end function
    integer(c_int) function lesgo_gpu_host_unregister(ctx, host) bind(c, name='lesgo_gpu_host_unregister')
        import
        type(c_ptr), value :: ctx
        type(c_ptr), value :: host
    end function
    integer(c_int) function lesgo_gpu_comm_p2p_import_null(ctx, nothing) bind(c, name='lesgo_gpu_comm_p2p_import')
        import
        type(c_ptr), value :: ctx
        type(c_ptr), value :: nothing          ! c_null_ptr: back to the NCCL all-to-alls
    end function
    integer(c_int) function lesgo_gpu_wavenumbers(ctx, kx, ky, k2) bind(c, name='lesgo_gpu_wavenumbers')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(out) :: kx(*), ky(*), k2(*)
    end function
    integer(c_int) function lesgo_gpu_padd(ctx, u_big, u, nplanes) bind(c, name='lesgo_gpu_padd')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(out) :: u_big(*)
        real(c_double), intent(in) :: u(*)
        integer(c_int), value :: nplanes
    end function
    integer(c_int) function lesgo_gpu_unpadd(ctx, cc, cc_big, nplanes) bind(c, name='lesgo_gpu_unpadd')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(out) :: cc(*)
        real(c_double), intent(in) :: cc_big(*)
        integer(c_int), value :: nplanes
    end function
    integer(c_int) function lesgo_gpu_fft_r2c(ctx, a, b, nplanes, big) bind(c, name='lesgo_gpu_fft_r2c')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(in) :: a(*)
        real(c_double), intent(inout) :: b(*)
        integer(c_int), value :: nplanes, big
    end function
    integer(c_int) function lesgo_gpu_fft_c2r(ctx, a, b, nplanes, big) bind(c, name='lesgo_gpu_fft_c2r')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(in) :: a(*)
        real(c_double), intent(inout) :: b(*)
        integer(c_int), value :: nplanes, big
    end function
    integer(c_int) function lesgo_gpu_ddx(ctx, f, dfdx) bind(c, name='lesgo_gpu_ddx')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(in) :: f(*)
        real(c_double), intent(inout) :: dfdx(*)
    end function
    integer(c_int) function lesgo_gpu_ddy(ctx, f, dfdy) bind(c, name='lesgo_gpu_ddy')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(in) :: f(*)
        real(c_double), intent(inout) :: dfdy(*)
    end function
    integer(c_int) function lesgo_gpu_ddxy(ctx, f, dfdx, dfdy) bind(c, name='lesgo_gpu_ddxy')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(in) :: f(*)
        real(c_double), intent(inout) :: dfdx(*), dfdy(*)
    end function
    integer(c_int) function lesgo_gpu_filt_da(ctx, f, dfdx, dfdy) bind(c, name='lesgo_gpu_filt_da')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(inout) :: f(*), dfdx(*), dfdy(*)
    end function
    integer(c_int) function lesgo_gpu_ddz_uv(ctx, f, dfdz) bind(c, name='lesgo_gpu_ddz_uv')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(in) :: f(*)
        real(c_double), intent(inout) :: dfdz(*)
    end function
    integer(c_int) function lesgo_gpu_ddz_w(ctx, f, dfdz) bind(c, name='lesgo_gpu_ddz_w')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(in) :: f(*)
        real(c_double), intent(inout) :: dfdz(*)
    end function
    integer(c_int) function lesgo_gpu_convec(ctx, u, v, w, dudy, dudz, dvdx, dvdz, dwdx, dwdy,      &
        RHSx, RHSy, RHSz) bind(c, name='lesgo_gpu_convec')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(in) :: u(*), v(*), w(*), dudy(*), dudz(*), dvdx(*), dvdz(*), dwdx(*), dwdy(*)
        real(c_double), intent(inout) :: RHSx(*), RHSy(*), RHSz(*)
    end function
    integer(c_int) function lesgo_gpu_press_stag_array(ctx, u, v, w, divtz, dt, tadv1, p, dpdx,      &
        dpdy, dpdz) bind(c, name='lesgo_gpu_press_stag_array')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(in) :: u(*), v(*), w(*), divtz(*)
        real(c_double), value :: dt, tadv1
        real(c_double), intent(inout) :: p(*), dpdx(*), dpdy(*), dpdz(*)
    end function
    integer(c_int) function lesgo_gpu_tridag_array(ctx, a, b, c, r, u, n) bind(c, name='lesgo_gpu_tridag_array')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(in) :: a(*), b(*), c(*), r(*)
        real(c_double), intent(inout) :: u(*)
        integer(c_int), value :: n
    end function
    !> test_filtermodule.f90:126-168: nplanes planes of f filtered in place with the real kernel G(lh, ny)
    integer(c_int) function lesgo_gpu_test_filter(ctx, f, G, nplanes) bind(c, name='lesgo_gpu_test_filter')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(inout) :: f(*)
        real(c_double), intent(in) :: G(*)
        integer(c_int), value :: nplanes
    end function
    !> mpi_defs.f90:167-264 on a (ld, ny, 0:nz) array: isync = MPI_SYNC_DOWN (1), MPI_SYNC_UP (2), MPI_SYNC_DOWNUP (3)
    integer(c_int) function lesgo_gpu_sync_real_array(ctx, var, isync) bind(c, name='lesgo_gpu_sync_real_array')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx
        real(c_double), intent(inout) :: var(*)
        integer(c_int), value :: isync
    end function
end interface

contains

!> Create the device context on first use (after read_input_conf and initialize_mpi):
!> one rank = one GPU.  The device is the NODE-LOCAL rank (MPI_Comm_split_type over shared
!> memory) modulo the device count, so `mpirun -np 8 lesgo-mpi` on one 8-GPU box lands every rank on its
!> own GPU without a CUDA_VISIBLE_DEVICES wrapper.  The NCCL id is made on coord 0 and broadcast with the
!> MPI communicator LESGO already owns (mpi_defs.f90:77-87).
!> The library addresses every field as (ld, ny, 0:nz), the layout of the MPI build (param.f90:68-72):
!> a serial build has lbz = 1 and would be shifted by one plane, so it is refused here.
subroutine gpu_require()
use param, only : nx, ny, nz, nz_tot, nproc, coord, L_x, L_y, dz, lbc_mom, ubc_mom, sgs, lbz
use messages, only : error
#ifdef PPMPI
use param, only : comm, ierr
use mpi
#endif
type(lesgo_gpu_dims) :: d
character(kind=c_char) :: id(128), blob(128), blobs(128 * 8)
integer :: device, local_comm, local_rank, p2p_ok, p2p_all
if (c_associated(gpu_ctx)) return
#ifndef PPMPI
call error('lesgo_gpu.gpu_require', 'liblesgo_cuda needs the MPI build of LESGO (USE_MPI, lbz = 0); run it with -np 1 for one GPU')
#endif
if (lbz /= 0) call error('lesgo_gpu.gpu_require', 'lbz must be 0 (MPI build): the library addresses fields as (ld, ny, 0:nz)')
device = -1
#ifdef PPMPI
call mpi_comm_split_type(comm, MPI_COMM_TYPE_SHARED, 0, MPI_INFO_NULL, local_comm, ierr)
call mpi_comm_rank(local_comm, local_rank, ierr)
call mpi_comm_free(local_comm, ierr)
device = -2 - local_rank        ! lesgo_gpu_create maps -2 - r to device mod(r, device count)
#endif
d = lesgo_gpu_dims(nx, ny, nz, nz_tot, nproc, coord, L_x, L_y, dz, lbc_mom, ubc_mom,             &
    merge(1, 0, sgs), device)
call gpu_check(lesgo_gpu_create(d, gpu_ctx), 'lesgo_gpu_create')
#ifdef PPMPI
if (nproc > 1) then
    if (coord == 0) call gpu_check(lesgo_gpu_comm_unique_id(id), 'lesgo_gpu_comm_unique_id')
    call mpi_bcast(id, 128, MPI_CHARACTER, 0, comm, ierr)
    call gpu_check(lesgo_gpu_comm_init(gpu_ctx, id), 'lesgo_gpu_comm_init')
    ! Pressure transposes over NVLink peer memory (one node, <= 8 ranks).  Optional and COLLECTIVE: a rank that
    ! cannot export or map a peer buffer (ranks on several nodes, GPUs without P2P) makes every rank stay on the
    ! NCCL all-to-alls -- the decision is all-reduced so the ranks never disagree.
    if (nproc <= 8) then
        p2p_ok = merge(1, 0, lesgo_gpu_comm_p2p_export(gpu_ctx, blob) == 0)
        call mpi_allreduce(p2p_ok, p2p_all, 1, MPI_INTEGER, MPI_MIN, comm, ierr)
        if (p2p_all == 1) then
            call mpi_allgather(blob, 128, MPI_CHARACTER, blobs, 128, MPI_CHARACTER, comm, ierr)
            p2p_ok = merge(1, 0, lesgo_gpu_comm_p2p_import(gpu_ctx, blobs) == 0)
            call mpi_allreduce(p2p_ok, p2p_all, 1, MPI_INTEGER, MPI_MIN, comm, ierr)
        end if
        if (p2p_all /= 1) call gpu_check(lesgo_gpu_comm_p2p_import_null(gpu_ctx, c_null_ptr), 'lesgo_gpu_comm_p2p_import')
    end if
end if
#endif
end subroutine gpu_require

!> Page-lock a module array that is handed to the per-routine entry points every step (sim_param.f90:52-82
!> allocatables are pageable): call once after sim_param_init, e.g. `call gpu_pin(u)`.
subroutine gpu_pin(a)
real(c_double), dimension(:,:,:), contiguous, target, intent(in) :: a
call gpu_require()
call gpu_check(lesgo_gpu_host_register(gpu_ctx, c_loc(a), int(size(a), c_size_t) * 8_c_size_t), 'lesgo_gpu_host_register')
end subroutine gpu_pin

!> Pin the arrays of the hot path (main.f90:161-172, 207, 317): called by the replacement init_fft.
subroutine gpu_pin_sim_param()
use sim_param
call gpu_pin(u); call gpu_pin(v); call gpu_pin(w)
call gpu_pin(dudx); call gpu_pin(dudy); call gpu_pin(dudz)
call gpu_pin(dvdx); call gpu_pin(dvdy); call gpu_pin(dvdz)
call gpu_pin(dwdx); call gpu_pin(dwdy); call gpu_pin(dwdz)
call gpu_pin(RHSx); call gpu_pin(RHSy); call gpu_pin(RHSz)
call gpu_pin(p); call gpu_pin(dpdx); call gpu_pin(dpdy); call gpu_pin(dpdz); call gpu_pin(divtz)
end subroutine gpu_pin_sim_param

!> test_filter / test_test_filter of test_filtermodule.f90:126-168 for ONE plane, with the module's own kernel:
!>     call gpu_test_filter(f, G_test)       replaces the body of test_filter(f)
!>     call gpu_test_filter(f, G_test_test)  replaces the body of test_test_filter(f)
subroutine gpu_test_filter(f, G)
real(c_double), dimension(:,:), intent(inout) :: f       !< (ld, ny)
real(c_double), dimension(:,:), intent(in) :: G          !< (lh, ny)
call gpu_require()
call gpu_check(lesgo_gpu_test_filter(gpu_ctx, f, G, 1_c_int), 'lesgo_gpu_test_filter')
end subroutine gpu_test_filter

!> mpi_sync_real_array (mpi_defs.f90:167-264) for a field the host holds, over the library's transport (NCCL or the
!> peer mappings) instead of MPI: same isync values as MPI_SYNC_DOWN / MPI_SYNC_UP / MPI_SYNC_DOWNUP.
subroutine gpu_sync_real_array(var, isync)
real(c_double), intent(inout) :: var(*)
integer, intent(in) :: isync
call gpu_require()
call gpu_check(lesgo_gpu_sync_real_array(gpu_ctx, var, int(isync, c_int)), 'lesgo_gpu_sync_real_array')
end subroutine gpu_sync_real_array

!> Non-zero return code -> the reference's fatal-error path (messages.f90:228-240).
subroutine gpu_check(rc, where)
use messages, only : error
integer(c_int), intent(in) :: rc
character(*), intent(in) :: where
character(kind=c_char), pointer :: msg(:)
character(512) :: text
integer :: i
if (rc == 0) return
text = ''
call c_f_pointer(lesgo_gpu_last_error(gpu_ctx), msg, [512])
do i = 1, 512
    if (msg(i) == c_null_char) exit
    text(i:i) = msg(i)
end do
call error('lesgo_gpu.'//where, trim(text))
end subroutine gpu_check

end module lesgo_gpu_mod
