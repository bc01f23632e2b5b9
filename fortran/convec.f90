!> Drop-in replacement for the external subroutine convec (reference convec.f90:21-334):
!> no arguments, works on the sim_param module arrays.
subroutine convec
use sim_param, only : u, v, w, dudy, dudz, dvdx, dvdz, dwdx, dwdy, RHSx, RHSy, RHSz
use lesgo_gpu_mod
implicit none
call gpu_require()
call gpu_check(lesgo_gpu_convec(gpu_ctx, u, v, w, dudy, dudz, dvdx, dvdz, dwdx, dwdy,             &
    RHSx, RHSy, RHSz), 'convec')
end subroutine convec
