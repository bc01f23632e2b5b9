!> Drop-in replacement for module fft (reference fft.f90:24-162): same public names.
!> init_fft makes its four plans exactly as the reference does (fft.f90:114-121) -- but the symbols
!> dfftw_plan_dft_r2c_2d / dfftw_plan_dft_c2r_2d / dfftw_execute_dft_r2c / dfftw_execute_dft_c2r /
!> dfftw_destroy_plan now resolve to liblesgo_cuda.so (include/lesgo_gpu.h, csrc/fftw_shim.cu), so callers
!> outside the hot path that still use the legacy FFTW API with these handles or with plans of their own
!> (test_filtermodule.f90:138,144; scalars.f90:513-624; turbine_indicator.f90:130-151) keep working and
!> libfftw3 is no longer linked.
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
integer*8 :: forw, back, forw_big, back_big
contains

subroutine init_fft()
use param, only : nx, ny, nx2, ny2
real(rprec), allocatable, dimension(:,:) :: data, data_big
integer, parameter :: FFTW_PATIENT = 32, FFTW_UNALIGNED = 2     ! fftw3.f; ignored by the library
call gpu_require()
allocate(data(ld, ny), data_big(ld_big, ny2))
call dfftw_plan_dft_r2c_2d(forw, nx, ny, data, data, FFTW_PATIENT, FFTW_UNALIGNED)
call dfftw_plan_dft_c2r_2d(back, nx, ny, data, data, FFTW_PATIENT, FFTW_UNALIGNED)
call dfftw_plan_dft_r2c_2d(forw_big, nx2, ny2, data_big, data_big, FFTW_PATIENT, FFTW_UNALIGNED)
call dfftw_plan_dft_c2r_2d(back_big, nx2, ny2, data_big, data_big, FFTW_PATIENT, FFTW_UNALIGNED)
deallocate(data, data_big)
allocate(kx(lh, ny), ky(lh, ny), k2(lh, ny))
call gpu_check(lesgo_gpu_wavenumbers(gpu_ctx, kx, ky, k2), 'init_fft')
call gpu_pin_sim_param()       ! sim_param_init ran before init_fft (initialize.f90:126,172)
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
