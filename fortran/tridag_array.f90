!> Drop-in replacement for tridag_array (reference tridag_array.f90:22-32, explicit-shape
!> arguments, symbol tridag_array_).  press_stag_array no longer calls it (the solve is
!> fused on the device), it is kept for any other caller.
subroutine tridag_array(a, b, c, r, u)
use types, only : rprec
use param, only : lh, ld, ny, nz
use lesgo_gpu_mod
implicit none
real(rprec), dimension(lh, ny, nz+1), intent(in) :: a, b, c
real(rprec), dimension(ld, ny, nz+1), intent(in) :: r
real(rprec), dimension(ld, ny, nz+1), intent(out) :: u
call gpu_require()
call gpu_check(lesgo_gpu_tridag_array(gpu_ctx, a, b, c, r, u, nz+1), 'tridag_array')
end subroutine tridag_array
