"""Op-stack table AIR — restates triton-air/src/table/op_stack.rs:33-205."""
from ..columns import MAIN, AUX, CH, Env

C, A = MAIN["op_stack"], AUX["op_stack"]
PADDING_VALUE = 2          # op_stack.rs:23
NUM_OP_STACK_REGISTERS = 16


def initial(b):
    e = Env(b)
    initial_stack_length = e.constant(NUM_OP_STACK_REGISTERS)
    padding_indicator = e.constant(PADDING_VALUE)
    stack_pointer_is_16 = e.main(C.StackPointer) - initial_stack_length
    compressed_row = (e.challenge(CH.OpStackClkWeight) * e.main(C.CLK)
                      + e.challenge(CH.OpStackIb1Weight) * e.main(C.IB1ShrinkStack)
                      + e.challenge(CH.OpStackPointerWeight) * initial_stack_length
                      + e.challenge(CH.OpStackFirstUnderflowElementWeight) * e.main(C.FirstUnderflowElement))
    rppa_initial = e.challenge(CH.OpStackIndeterminate) - compressed_row
    rppa_has_accumulated_first_row = e.aux(A.RunningProductPermArg) - rppa_initial
    rppa_is_default_initial = e.aux(A.RunningProductPermArg) - e.x_constant(1)
    first_row_is_padding_row = e.main(C.IB1ShrinkStack) - padding_indicator
    first_row_is_not_padding_row = e.main(C.IB1ShrinkStack) * (e.main(C.IB1ShrinkStack) - e.constant(1))
    rppa_starts_correctly = (rppa_has_accumulated_first_row * first_row_is_padding_row
                             + rppa_is_default_initial * first_row_is_not_padding_row)
    lookup_argument_initial = e.x_constant(0)
    cjd_init = e.aux(A.ClockJumpDifferenceLookupClientLogDerivative) - lookup_argument_initial
    return [stack_pointer_is_16, rppa_starts_correctly, cjd_init]


def consistency(b):
    e = Env(b)
    ib1 = lambda: e.main(C.IB1ShrinkStack)
    ib1_is_legal = ib1() * (ib1() - e.constant(1)) * (ib1() - e.constant(PADDING_VALUE))
    return [ib1_is_legal]


def transition(b):
    e = Env(b)
    one = e.constant(1)
    padding_indicator = e.constant(PADDING_VALUE)

    clk = e.cur_main(C.CLK)
    ib1_shrink_stack = e.cur_main(C.IB1ShrinkStack)
    stack_pointer = e.cur_main(C.StackPointer)
    first_underflow_element = e.cur_main(C.FirstUnderflowElement)
    rppa = e.cur_aux(A.RunningProductPermArg)
    cjd = e.cur_aux(A.ClockJumpDifferenceLookupClientLogDerivative)

    clk_next = e.next_main(C.CLK)
    ib1_shrink_stack_next = e.next_main(C.IB1ShrinkStack)
    stack_pointer_next = e.next_main(C.StackPointer)
    first_underflow_element_next = e.next_main(C.FirstUnderflowElement)
    rppa_next = e.next_aux(A.RunningProductPermArg)
    cjd_next = e.next_aux(A.ClockJumpDifferenceLookupClientLogDerivative)

    c0 = (stack_pointer_next - stack_pointer - one) * (stack_pointer_next - stack_pointer)
    c1 = ((stack_pointer_next - stack_pointer - one)
          * (first_underflow_element_next - first_underflow_element) * ib1_shrink_stack_next)

    next_row_is_padding_row = ib1_shrink_stack_next - padding_indicator
    c2 = ib1_shrink_stack * (ib1_shrink_stack - one) * next_row_is_padding_row

    compressed_row = (e.challenge(CH.OpStackClkWeight) * clk_next
                      + e.challenge(CH.OpStackIb1Weight) * ib1_shrink_stack_next
                      + e.challenge(CH.OpStackPointerWeight) * stack_pointer_next
                      + e.challenge(CH.OpStackFirstUnderflowElementWeight) * first_underflow_element_next)
    rppa_updates = rppa_next - rppa * (e.challenge(CH.OpStackIndeterminate) - compressed_row)
    next_row_is_not_padding_row = ib1_shrink_stack_next * (ib1_shrink_stack_next - one)
    rppa_remains = rppa_next - rppa
    rppa_updates_correctly = rppa_updates * next_row_is_padding_row + rppa_remains * next_row_is_not_padding_row

    clk_diff = clk_next - clk
    log_derivative_accumulates = ((cjd_next - cjd)
                                  * (e.challenge(CH.ClockJumpDifferenceLookupIndeterminate) - clk_diff) - one)
    log_derivative_remains = cjd_next - cjd
    t0 = log_derivative_accumulates * (stack_pointer_next - stack_pointer - one) * next_row_is_padding_row
    t1 = log_derivative_remains * (stack_pointer_next - stack_pointer)
    t2 = log_derivative_remains * next_row_is_not_padding_row
    log_derivative_updates_correctly = t0 + t1 + t2
    return [c0, c1, c2, rppa_updates_correctly, log_derivative_updates_correctly]


def terminal(b):
    return []
