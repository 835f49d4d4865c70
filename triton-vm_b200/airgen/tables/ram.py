"""RAM table AIR — restates triton-air/src/table/ram.rs:29-281."""
from ..columns import MAIN, AUX, CH, Env

C, A = MAIN["ram"], AUX["ram"]
INSTRUCTION_TYPE_WRITE, INSTRUCTION_TYPE_READ, PADDING_INDICATOR = 0, 1, 2   # ram.rs:20-22


def initial(b):
    e = Env(b)
    first_row_is_padding_row = e.main(C.InstructionType) - e.constant(PADDING_INDICATOR)
    first_row_is_not_padding_row = ((e.main(C.InstructionType) - e.constant(INSTRUCTION_TYPE_READ))
                                    * (e.main(C.InstructionType) - e.constant(INSTRUCTION_TYPE_WRITE)))
    bcpc0_is_0 = e.main(C.BezoutCoefficientPolynomialCoefficient0)
    bc0_is_0 = e.aux(A.BezoutCoefficient0)
    bc1_is_bcpc1 = e.aux(A.BezoutCoefficient1) - e.main(C.BezoutCoefficientPolynomialCoefficient1)
    formal_derivative_is_1 = e.aux(A.FormalDerivative) - e.constant(1)
    rp_init = (e.aux(A.RunningProductOfRAMP) - e.challenge(CH.RamTableBezoutRelationIndeterminate)
               + e.main(C.RamPointer))
    cjd_default = e.aux(A.ClockJumpDifferenceLookupClientLogDerivative) - e.x_constant(0)
    compressed_row = (e.main(C.CLK) * e.challenge(CH.RamClkWeight)
                      + e.main(C.InstructionType) * e.challenge(CH.RamInstructionTypeWeight)
                      + e.main(C.RamPointer) * e.challenge(CH.RamPointerWeight)
                      + e.main(C.RamValue) * e.challenge(CH.RamValueWeight))
    rppa_accumulated = e.aux(A.RunningProductPermArg) - e.challenge(CH.RamIndeterminate) + compressed_row
    rppa_default = e.aux(A.RunningProductPermArg) - e.x_constant(1)
    rppa_starts_correctly = rppa_accumulated * first_row_is_padding_row + rppa_default * first_row_is_not_padding_row
    return [bcpc0_is_0, bc0_is_0, bc1_is_bcpc1, rp_init, formal_derivative_is_1, rppa_starts_correctly, cjd_default]


def consistency(b):
    e = Env(b)
    it = lambda: e.main(C.InstructionType)
    legal = ((it() - e.constant(INSTRUCTION_TYPE_WRITE)) * (it() - e.constant(INSTRUCTION_TYPE_READ))
             * (it() - e.constant(PADDING_INDICATOR)))
    return [legal]


def transition(b):
    e = Env(b)
    one = e.constant(1)
    bezout_challenge = e.challenge(CH.RamTableBezoutRelationIndeterminate)

    clock = e.cur_main(C.CLK)
    ram_pointer = e.cur_main(C.RamPointer)
    ram_value = e.cur_main(C.RamValue)
    instruction_type = e.cur_main(C.InstructionType)
    iord = e.cur_main(C.InverseOfRampDifference)
    bcpc0 = e.cur_main(C.BezoutCoefficientPolynomialCoefficient0)
    bcpc1 = e.cur_main(C.BezoutCoefficientPolynomialCoefficient1)

    rp = e.cur_aux(A.RunningProductOfRAMP)
    fd = e.cur_aux(A.FormalDerivative)
    bc0 = e.cur_aux(A.BezoutCoefficient0)
    bc1 = e.cur_aux(A.BezoutCoefficient1)
    rppa = e.cur_aux(A.RunningProductPermArg)
    cjd = e.cur_aux(A.ClockJumpDifferenceLookupClientLogDerivative)

    clock_next = e.next_main(C.CLK)
    ram_pointer_next = e.next_main(C.RamPointer)
    ram_value_next = e.next_main(C.RamValue)
    instruction_type_next = e.next_main(C.InstructionType)
    bcpc0_next = e.next_main(C.BezoutCoefficientPolynomialCoefficient0)
    bcpc1_next = e.next_main(C.BezoutCoefficientPolynomialCoefficient1)

    rp_next = e.next_aux(A.RunningProductOfRAMP)
    fd_next = e.next_aux(A.FormalDerivative)
    bc0_next = e.next_aux(A.BezoutCoefficient0)
    bc1_next = e.next_aux(A.BezoutCoefficient1)
    rppa_next = e.next_aux(A.RunningProductPermArg)
    cjd_next = e.next_aux(A.ClockJumpDifferenceLookupClientLogDerivative)

    next_row_is_padding_row = instruction_type_next - e.constant(PADDING_INDICATOR)
    c0 = ((instruction_type - e.constant(INSTRUCTION_TYPE_READ))
          * (instruction_type - e.constant(INSTRUCTION_TYPE_WRITE)) * next_row_is_padding_row)

    ram_pointer_difference = ram_pointer_next - ram_pointer
    ram_pointer_changes = one - ram_pointer_difference * iord

    c1 = iord * ram_pointer_changes
    c2 = ram_pointer_difference * ram_pointer_changes
    c3 = (ram_pointer_changes * (e.constant(INSTRUCTION_TYPE_WRITE) - instruction_type_next)
          * (ram_value_next - ram_value))
    c4 = ram_pointer_changes * (bcpc0_next - bcpc0)
    c5 = ram_pointer_changes * (bcpc1_next - bcpc1)
    c6 = (ram_pointer_difference * (rp_next - rp * (bezout_challenge - ram_pointer_next))
          + ram_pointer_changes * (rp_next - rp))
    c7 = (ram_pointer_difference * (fd_next - rp - (bezout_challenge - ram_pointer_next) * fd)
          + ram_pointer_changes * (fd_next - fd))
    c8 = (ram_pointer_difference * (bc0_next - bezout_challenge * bc0 - bcpc0_next)
          + ram_pointer_changes * (bc0_next - bc0))
    c9 = (ram_pointer_difference * (bc1_next - bezout_challenge * bc1 - bcpc1_next)
          + ram_pointer_changes * (bc1_next - bc1))

    compressed_row = (clock_next * e.challenge(CH.RamClkWeight)
                      + ram_pointer_next * e.challenge(CH.RamPointerWeight)
                      + ram_value_next * e.challenge(CH.RamValueWeight)
                      + instruction_type_next * e.challenge(CH.RamInstructionTypeWeight))
    rppa_accumulates_next_row = rppa_next - rppa * (e.challenge(CH.RamIndeterminate) - compressed_row)
    next_row_is_not_padding_row = ((instruction_type_next - e.constant(INSTRUCTION_TYPE_READ))
                                   * (instruction_type_next - e.constant(INSTRUCTION_TYPE_WRITE)))
    rppa_remains_unchanged = rppa_next - rppa
    c10 = rppa_accumulates_next_row * next_row_is_padding_row + rppa_remains_unchanged * next_row_is_not_padding_row

    clock_difference = clock_next - clock
    ld_accumulates = ((cjd_next - cjd)
                      * (e.challenge(CH.ClockJumpDifferenceLookupIndeterminate) - clock_difference) - one)
    ld_remains = cjd_next - cjd
    t0 = ld_accumulates * ram_pointer_changes * next_row_is_padding_row
    t1 = ld_remains * ram_pointer_difference * next_row_is_padding_row
    t2 = ld_remains * next_row_is_not_padding_row
    c11 = t0 + t1 + t2
    return [c0, c1, c2, c3, c4, c5, c6, c7, c8, c9, c10, c11]


def terminal(b):
    e = Env(b)
    bezout_relation_holds = (e.aux(A.BezoutCoefficient0) * e.aux(A.RunningProductOfRAMP)
                             + e.aux(A.BezoutCoefficient1) * e.aux(A.FormalDerivative) - e.constant(1))
    return [bezout_relation_holds]
