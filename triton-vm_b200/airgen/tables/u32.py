"""U32 table AIR — restates triton-air/src/table/u32.rs:27-389."""
from ..columns import MAIN, AUX, CH, Env
from ..isa import OPCODE

C, A = MAIN["u32"], AUX["u32"]


def instruction_deselector(e, instruction_to_select, current_instruction):
    """u32.rs:371-389: fold(b_constant(1), mul) over (ci - opcode) of the other u32 instructions."""
    acc = e.constant(1)
    for instr in ["split", "lt", "and", "log_2_floor", "pow", "pop_count"]:
        if instr == instruction_to_select:
            continue
        acc = acc * (current_instruction - e.constant(OPCODE[instr]))
    return acc


def initial(b):
    e = Env(b)
    one = e.constant(1)
    copy_flag = e.main(C.CopyFlag)
    lhs = e.main(C.LHS)
    rhs = e.main(C.RHS)
    ci = e.main(C.CI)
    result = e.main(C.Result)
    lookup_multiplicity = e.main(C.LookupMultiplicity)
    rsld = e.aux(A.LookupServerLogDerivative)
    compressed_row = (e.challenge(CH.U32LhsWeight) * lhs + e.challenge(CH.U32RhsWeight) * rhs
                      + e.challenge(CH.U32CiWeight) * ci + e.challenge(CH.U32ResultWeight) * result)
    if_1 = copy_flag * (rsld * (e.challenge(CH.U32Indeterminate) - compressed_row) - lookup_multiplicity)
    default_initial = e.x_constant(0)
    if_0 = (copy_flag - one) * (rsld - default_initial)
    return [if_0 + if_1]


def consistency(b):
    e = Env(b)
    one = lambda: e.constant(1)
    two = lambda: e.constant(2)
    copy_flag = e.main(C.CopyFlag)
    bits = e.main(C.Bits)
    bits_minus_33_inv = e.main(C.BitsMinus33Inv)
    ci = e.main(C.CI)
    lhs = e.main(C.LHS)
    lhs_inv = e.main(C.LhsInv)
    rhs = e.main(C.RHS)
    rhs_inv = e.main(C.RhsInv)
    result = e.main(C.Result)
    lookup_multiplicity = e.main(C.LookupMultiplicity)
    desel = lambda instr: instruction_deselector(e, instr, ci)

    copy_flag_is_bit = copy_flag * (one() - copy_flag)
    copy_flag_is_0_or_bits_is_0 = copy_flag * bits
    bits_minus_33_inv_ok = one() - bits_minus_33_inv * (bits - e.constant(33))
    lhs_inv_ok = lhs_inv * (one() - lhs * lhs_inv)
    lhs_ok = lhs * (one() - lhs * lhs_inv)
    rhs_inv_ok = rhs_inv * (one() - rhs * rhs_inv)
    rhs_ok = rhs * (one() - rhs * rhs_inv)
    lt0 = (desel("lt") * (copy_flag - one()) * (one() - lhs * lhs_inv) * (one() - rhs * rhs_inv)
           * (result - two()))
    lt1 = desel("lt") * copy_flag * (one() - lhs * lhs_inv) * (one() - rhs * rhs_inv) * result
    and_ = desel("and") * (one() - lhs * lhs_inv) * (one() - rhs * rhs_inv) * result
    pow_ = desel("pow") * (one() - rhs * rhs_inv) * (result - one())
    log2 = desel("log_2_floor") * (copy_flag - one()) * (one() - lhs * lhs_inv) * (result + one())
    popc = desel("pop_count") * (one() - lhs * lhs_inv) * result
    crash = desel("log_2_floor") * copy_flag * (one() - lhs * lhs_inv)
    mult0 = (copy_flag - one()) * lookup_multiplicity
    return [copy_flag_is_bit, copy_flag_is_0_or_bits_is_0, bits_minus_33_inv_ok, lhs_inv_ok, lhs_ok, rhs_inv_ok, rhs_ok,
            lt0, lt1, and_, pow_, log2, popc, crash, mult0]


def transition(b):
    e = Env(b)
    one = lambda: e.constant(1)
    two = lambda: e.constant(2)

    copy_flag = e.cur_main(C.CopyFlag)
    bits = e.cur_main(C.Bits)
    ci = e.cur_main(C.CI)
    lhs = e.cur_main(C.LHS)
    rhs = e.cur_main(C.RHS)
    result = e.cur_main(C.Result)
    rsld = e.cur_aux(A.LookupServerLogDerivative)

    copy_flag_next = e.next_main(C.CopyFlag)
    bits_next = e.next_main(C.Bits)
    ci_next = e.next_main(C.CI)
    lhs_next = e.next_main(C.LHS)
    rhs_next = e.next_main(C.RHS)
    result_next = e.next_main(C.Result)
    lhs_inv_next = e.next_main(C.LhsInv)
    lookup_multiplicity_next = e.next_main(C.LookupMultiplicity)
    rsld_next = e.next_aux(A.LookupServerLogDerivative)

    desel = lambda instr: instruction_deselector(e, instr, ci_next)

    ci_is_pow = ci - e.constant(OPCODE["pow"])
    lhs_lsb = lhs - two() * lhs_next
    rhs_lsb = rhs - two() * rhs_next

    c0 = copy_flag_next * lhs * ci_is_pow
    c1 = copy_flag_next * rhs
    c2 = (copy_flag_next - one()) * (ci_next - ci)
    c3 = (copy_flag_next - one()) * lhs * ci_is_pow * (bits_next - bits - one())
    c4 = (copy_flag_next - one()) * rhs * (bits_next - bits - one())
    c5 = (copy_flag_next - one()) * ci_is_pow * lhs_lsb * (lhs_lsb - one())
    c6 = (copy_flag_next - one()) * rhs_lsb * (rhs_lsb - one())

    c7 = (copy_flag_next - one()) * desel("lt") * (result_next - one()) * (result_next - two()) * result
    c8 = (copy_flag_next - one()) * desel("lt") * result_next * (result_next - two()) * (result - one())
    c9 = ((copy_flag_next - one()) * desel("lt") * result_next * (result_next - one())
          * (lhs_lsb - one()) * rhs_lsb * (result - one()))
    c10 = ((copy_flag_next - one()) * desel("lt") * result_next * (result_next - one())
           * lhs_lsb * (rhs_lsb - one()) * result)
    c11 = ((copy_flag_next - one()) * desel("lt") * result_next * (result_next - one())
           * (one() - lhs_lsb - rhs_lsb + two() * lhs_lsb * rhs_lsb)
           * (copy_flag - one()) * (result - two()))
    c12 = ((copy_flag_next - one()) * desel("lt") * result_next * (result_next - one())
           * (one() - lhs_lsb - rhs_lsb + two() * lhs_lsb * rhs_lsb)
           * copy_flag * result)

    c13 = (copy_flag_next - one()) * desel("and") * (result - two() * result_next - lhs_lsb * rhs_lsb)

    c14 = ((copy_flag_next - one()) * desel("log_2_floor") * (one() - lhs_next * lhs_inv_next) * lhs
           * (result - bits))
    c15 = (copy_flag_next - one()) * desel("log_2_floor") * lhs_next * (result_next - result)

    c16 = (copy_flag_next - one()) * desel("pow") * (lhs_next - lhs)
    c17 = (copy_flag_next - one()) * desel("pow") * (rhs_lsb - one()) * (result - result_next * result_next)
    c18 = (copy_flag_next - one()) * desel("pow") * rhs_lsb * (result - result_next * result_next * lhs)

    c19 = (copy_flag_next - one()) * desel("pop_count") * (result - result_next - lhs_lsb)

    c20 = (copy_flag_next - one()) * (rsld_next - rsld)

    compressed_row_next = (e.challenge(CH.U32CiWeight) * ci_next + e.challenge(CH.U32LhsWeight) * lhs_next
                           + e.challenge(CH.U32RhsWeight) * rhs_next + e.challenge(CH.U32ResultWeight) * result_next)
    c21 = copy_flag_next * ((rsld_next - rsld) * (e.challenge(CH.U32Indeterminate) - compressed_row_next)
                            - lookup_multiplicity_next)
    return [c0, c1, c2, c3, c4, c5, c6, c7, c8, c9, c10, c11, c12, c13, c14, c15, c16, c17, c18, c19, c20, c21]


def terminal(b):
    e = Env(b)
    ci = e.main(C.CI)
    lhs = e.main(C.LHS)
    rhs = e.main(C.RHS)
    lhs_is_0_or_ci_is_pow = lhs * (ci - e.constant(OPCODE["pow"]))
    return [lhs_is_0_or_ci_is_pow, rhs]
