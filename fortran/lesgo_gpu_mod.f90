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
end interface

contains

!> Create the device context on first use (after read_input_conf and initialize_mpi):
!> one rank = one GPU, device = local rank.  The NCCL id is made on coord 0 and
!> broadcast with the MPI communicator LESGO already owns (mpi_defs.f90:77-87).
subroutine gpu_require()
use param, only : nx, ny, nz, nz_tot, nproc, coord, L_x, L_y, dz, lbc_mom, ubc_mom, sgs
#ifdef PPMPI
use param, only : comm, ierr
use mpi
#endif
type(lesgo_gpu_dims) :: d
character(kind=c_char) :: id(128), blob(128), blobs(128 * 8)
if (c_associated(gpu_ctx)) return
d = lesgo_gpu_dims(nx, ny, nz, nz_tot, nproc, coord, L_x, L_y, dz, lbc_mom, ubc_mom,             &
    merge(1, 0, sgs), -1)
call gpu_check(lesgo_gpu_create(d, gpu_ctx), 'lesgo_gpu_create')
#ifdef PPMPI
if (nproc > 1) then
    if (coord == 0) call gpu_check(lesgo_gpu_comm_unique_id(id), 'lesgo_gpu_comm_unique_id')
    call mpi_bcast(id, 128, MPI_CHARACTER, 0, comm, ierr)
    call gpu_check(lesgo_gpu_comm_init(gpu_ctx, id), 'lesgo_gpu_comm_init')
    ! pressure transposes over NVLink peer memory (one node): gather every rank's 128-byte export
    if (nproc <= 8) then
        call gpu_check(lesgo_gpu_comm_p2p_export(gpu_ctx, blob), 'lesgo_gpu_comm_p2p_export')
        call mpi_allgather(blob, 128, MPI_CHARACTER, blobs, 128, MPI_CHARACTER, comm, ierr)
        call gpu_check(lesgo_gpu_comm_p2p_import(gpu_ctx, blobs), 'lesgo_gpu_comm_p2p_import')
    end if
end if
#endif
end subroutine gpu_require

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
