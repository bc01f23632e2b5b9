!> Drop-in replacement for module derivatives (reference derivatives.f90:21-315): same
!> public names and assumed-shape interfaces, bodies forward to liblesgo_cuda.so.
!> Assumed-shape dummies `dimension(:,:,lbz:)` arrive by descriptor, which is why this
!> wrapper has to stay Fortran; the actual arguments in LESGO are whole contiguous module
!> arrays (ld, ny, lbz:nz), so passing them on as `f` gives C the base address.
module derivatives
use types, only : rprec
use lesgo_gpu_mod
implicit none
save
private
public ddx, ddy, ddxy, filt_da, ddz_uv, ddz_w
contains

subroutine ddx(f, dfdx, lbz)
integer, intent(in) :: lbz
real(rprec), dimension(:,:,lbz:), contiguous, intent(in) :: f
real(rprec), dimension(:,:,lbz:), contiguous, intent(inout) :: dfdx
call gpu_require()
call gpu_check(lesgo_gpu_ddx(gpu_ctx, f, dfdx), 'ddx')
end subroutine ddx

subroutine ddy(f, dfdy, lbz)
integer, intent(in) :: lbz
real(rprec), dimension(:,:,lbz:), contiguous, intent(in) :: f
real(rprec), dimension(:,:,lbz:), contiguous, intent(inout) :: dfdy
call gpu_require()
call gpu_check(lesgo_gpu_ddy(gpu_ctx, f, dfdy), 'ddy')
end subroutine ddy

subroutine ddxy(f, dfdx, dfdy, lbz)
integer, intent(in) :: lbz
real(rprec), dimension(:,:,lbz:), contiguous, intent(in) :: f
real(rprec), dimension(:,:,lbz:), contiguous, intent(inout) :: dfdx, dfdy
call gpu_require()
call gpu_check(lesgo_gpu_ddxy(gpu_ctx, f, dfdx, dfdy), 'ddxy')
end subroutine ddxy

subroutine filt_da(f, dfdx, dfdy, lbz)
integer, intent(in) :: lbz
real(rprec), dimension(:,:,lbz:), contiguous, intent(inout) :: f, dfdx, dfdy
call gpu_require()
call gpu_check(lesgo_gpu_filt_da(gpu_ctx, f, dfdx, dfdy), 'filt_da')
end subroutine filt_da

subroutine ddz_uv(f, dfdz, lbz)
integer, intent(in) :: lbz
real(rprec), dimension(:,:,lbz:), contiguous, intent(in) :: f
real(rprec), dimension(:,:,lbz:), contiguous, intent(inout) :: dfdz
call gpu_require()
call gpu_check(lesgo_gpu_ddz_uv(gpu_ctx, f, dfdz), 'ddz_uv')
end subroutine ddz_uv

subroutine ddz_w(f, dfdz, lbz)
integer, intent(in) :: lbz
real(rprec), dimension(:,:,lbz:), contiguous, intent(in) :: f
real(rprec), dimension(:,:,lbz:), contiguous, intent(inout) :: dfdz
call gpu_require()
call gpu_check(lesgo_gpu_ddz_w(gpu_ctx, f, dfdz), 'ddz_w')
end subroutine ddz_w

end module derivatives
