!> Drop-in replacement for module fft (reference fft.f90:24-162): same public names.
!> The four FFTW plan handles become tags understood by the dfftw_* shims at the bottom,
!> so callers outside the hot path that still use the legacy FFTW API with these plans
!> (test_filtermodule.f90:138,144; scalars.f90; turbine_indicator.f90:130-151) keep
!> working without FFTW being linked.
module fft
use types, only : rprec
use param, only : ld, lh, ny, ld_big, ny2
use lesgo_gpu_mod
implicit none
save
public :: padd, unpadd, init_fft
public :: kx, ky, k2
public :: forw, back, forw_big, back_big
real(rprec), allocatable, dimension(:,:) :: kx, ky, k2
integer*8 :: forw = 1, back = 2, forw_big = 3, back_big = 4
contains

subroutine init_fft()
call gpu_require()
allocate(kx(lh, ny), ky(lh, ny), k2(lh, ny))
call gpu_check(lesgo_gpu_wavenumbers(gpu_ctx, kx, ky, k2), 'init_fft')
end subroutine init_fft

subroutine padd(u_big, u)
real(rprec), dimension(ld, ny), intent(in) :: u
real(rprec), dimension(ld_big, ny2), intent(out) :: u_big
call gpu_check(lesgo_gpu_padd(gpu_ctx, u_big, u, 1), 'padd')
end subroutine padd

subroutine unpadd(cc, cc_big)
real(rprec), dimension(ld, ny) :: cc
real(rprec), dimension(ld_big, ny2) :: cc_big
call gpu_check(lesgo_gpu_unpadd(gpu_ctx, cc, cc_big, 1), 'unpadd')
end subroutine unpadd

end module fft

!> FFTW legacy-Fortran entry points for the four plans above (one plane per call).
subroutine dfftw_execute_dft_r2c(plan, a, b)
use lesgo_gpu_mod
implicit none
integer*8, intent(in) :: plan
real(c_double) :: a(*), b(*)
call gpu_check(lesgo_gpu_fft_r2c(gpu_ctx, a, b, 1, merge(1, 0, plan == 3)), 'dfftw_execute_dft_r2c')
end subroutine dfftw_execute_dft_r2c

subroutine dfftw_execute_dft_c2r(plan, a, b)
use lesgo_gpu_mod
implicit none
integer*8, intent(in) :: plan
real(c_double) :: a(*), b(*)
call gpu_check(lesgo_gpu_fft_c2r(gpu_ctx, a, b, 1, merge(1, 0, plan == 4)), 'dfftw_execute_dft_c2r')
end subroutine dfftw_execute_dft_c2r
