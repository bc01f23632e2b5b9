!> Drop-in replacement for press_stag_array (reference press_stag_array.f90:21-290).
!> p is (ld, ny, 0:nz); dpdx, dpdy, dpdz are declared 1:nz in sim_param.f90:70-72, the
!> library addresses every array from plane 0, so their plane-1 address is handed over
!> shifted down by one plane (plane 0 of those three is never touched by the library).
subroutine press_stag_array()
use types, only : rprec
use param, only : ld, ny, dt, tadv1
use sim_param, only : u, v, w, divtz, p, dpdx, dpdy, dpdz
use lesgo_gpu_mod
use iso_c_binding
implicit none
interface
    integer(c_int) function press_raw(ctx, u, v, w, divtz, dt, tadv1, p, dpdx, dpdy, dpdz)         &
        bind(c, name='lesgo_gpu_press_stag_array')
        import :: c_int, c_ptr, c_double
        type(c_ptr), value :: ctx, dpdx, dpdy, dpdz
        real(c_double), intent(in) :: u(*), v(*), w(*), divtz(*)
        real(c_double), value :: dt, tadv1
        real(c_double), intent(inout) :: p(*)
    end function
end interface
integer(c_intptr_t) :: shift
call gpu_require()
shift = int(ld, c_intptr_t) * int(ny, c_intptr_t) * 8_c_intptr_t
call gpu_check(press_raw(gpu_ctx, u, v, w, divtz, dt, tadv1, p,                                    &
    transfer(transfer(c_loc(dpdx), shift) - shift, c_null_ptr),                                    &
    transfer(transfer(c_loc(dpdy), shift) - shift, c_null_ptr),                                    &
    transfer(transfer(c_loc(dpdz), shift) - shift, c_null_ptr)), 'press_stag_array')
end subroutine press_stag_array
