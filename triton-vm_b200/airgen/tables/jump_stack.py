"""Jump-stack table AIR — restates triton-air/src/table/jump_stack.rs:43-163."""
from ..columns import MAIN, AUX, CH, Env
from ..isa import OPCODE

C, A = MAIN["jump_stack"], AUX["jump_stack"]


def initial(b):
    e = Env(b)
    clk = e.main(C.CLK)
    jsp = e.main(C.JSP)
    jso = e.main(C.JSO)
    jsd = e.main(C.JSD)
    ci = e.main(C.CI)
    rppa = e.aux(A.RunningProductPermArg)
    cjd = e.aux(A.ClockJumpDifferenceLookupClientLogDerivative)
    processor_perm_indeterminate = e.challenge(CH.JumpStackIndeterminate)
    compressed_row = e.challenge(CH.JumpStackCiWeight) * ci
    rppa_starts_correctly = rppa - (processor_perm_indeterminate - compressed_row)
    cjd_starts_correctly = cjd - e.x_constant(0)
    return [clk, jsp, jso, jsd, rppa_starts_correctly, cjd_starts_correctly]


def consistency(b):
    return []


def transition(b):
    e = Env(b)
    one = lambda: e.constant(1)
    call_opcode = e.constant(OPCODE["call"])
    return_opcode = e.constant(OPCODE["return"])
    recurse_or_return_opcode = e.constant(OPCODE["recurse_or_return"])

    clk = e.cur_main(C.CLK)
    ci = e.cur_main(C.CI)
    jsp = e.cur_main(C.JSP)
    jso = e.cur_main(C.JSO)
    jsd = e.cur_main(C.JSD)
    rppa = e.cur_aux(A.RunningProductPermArg)
    cjd = e.cur_aux(A.ClockJumpDifferenceLookupClientLogDerivative)

    clk_next = e.next_main(C.CLK)
    ci_next = e.next_main(C.CI)
    jsp_next = e.next_main(C.JSP)
    jso_next = e.next_main(C.JSO)
    jsd_next = e.next_main(C.JSD)
    rppa_next = e.next_aux(A.RunningProductPermArg)
    cjd_next = e.next_aux(A.ClockJumpDifferenceLookupClientLogDerivative)

    jsp_inc_or_stays = (jsp_next - jsp - one()) * (jsp_next - jsp)
    jsp_inc_by_one_or_ci_can_return = ((jsp_next - jsp - one()) * (ci - return_opcode)
                                       * (ci - recurse_or_return_opcode))
    c1 = jsp_inc_by_one_or_ci_can_return * (jso_next - jso)
    c2 = jsp_inc_by_one_or_ci_can_return * (jsd_next - jsd)
    c3 = jsp_inc_by_one_or_ci_can_return * (clk_next - clk - one()) * (ci - call_opcode)

    compressed_row = (e.challenge(CH.JumpStackClkWeight) * clk_next
                      + e.challenge(CH.JumpStackCiWeight) * ci_next
                      + e.challenge(CH.JumpStackJspWeight) * jsp_next
                      + e.challenge(CH.JumpStackJsoWeight) * jso_next
                      + e.challenge(CH.JumpStackJsdWeight) * jsd_next)
    rppa_updates_correctly = rppa_next - rppa * (e.challenge(CH.JumpStackIndeterminate) - compressed_row)

    log_derivative_remains = cjd_next - cjd
    clk_diff = clk_next - clk
    log_derivative_accumulates = ((cjd_next - cjd)
                                  * (e.challenge(CH.ClockJumpDifferenceLookupIndeterminate) - clk_diff) - one())
    log_derivative_updates_correctly = ((jsp_next - jsp - one()) * log_derivative_accumulates
                                        + (jsp_next - jsp) * log_derivative_remains)
    return [jsp_inc_or_stays, c1, c2, c3, rppa_updates_correctly, log_derivative_updates_correctly]


def terminal(b):
    return []
