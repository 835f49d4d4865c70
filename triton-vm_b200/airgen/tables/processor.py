"""Processor table AIR — restates triton-air/src/table/processor.rs:73-3134 function by function.

Evaluation order is kept identical to the Rust source (operands left to right, `fold`/`sum`
consuming lazily mapped iterators, eager array `.map`, eager `unwrap_or(&constant(0))`
arguments), because node ids — and through them the degree-lowering tie-breaks — depend on it.
"""
from ..circuit import msum
from ..columns import MAIN, AUX, CH, Env
from ..isa import ALL_INSTRUCTIONS, OPCODE, ib as instr_ib

C, A = MAIN["processor"], AUX["processor"]
NUM_OP_STACK_REGISTERS = 16
RATE = 10
LEGAL_NUMBER_OF_WORDS = [1, 2, 3, 4, 5]                      # NumberOfWords::legal_values
ILLEGAL_NUMBER_OF_WORDS = [0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]
RAM_INSTRUCTION_TYPE_WRITE, RAM_INSTRUCTION_TYPE_READ = 0, 1
P = (1 << 64) - (1 << 32) + 1


def st(i): return getattr(C, f"ST{i}")
def hvc(i): return getattr(C, f"HV{i}")
def ibc(i): return getattr(C, f"IB{i}")
def stack_weight(e, i): return e.challenge(getattr(CH, f"StackWeight{i}"))


# ---- initial / consistency / terminal (processor.rs:76-302, 351-361) -------------------
def instruction_deselector_common(e, instruction, ib_polys):
    """processor.rs:709-737"""
    one = lambda: e.constant(1)
    selector_bits = [instr_ib(instruction, k) for k in range(7)]
    acc = one()
    for x_ib, ib_of_instr in zip(ib_polys, selector_bits):
        term = x_ib * e.constant(ib_of_instr) + (one() - x_ib) * e.constant((1 - ib_of_instr) % P)
        acc = acc * term
    return acc


def instruction_deselector_current_row(e, instruction):
    return instruction_deselector_common(e, instruction, [e.cur_main(ibc(k)) for k in range(7)])


def instruction_deselector_next_row(e, instruction):
    return instruction_deselector_common(e, instruction, [e.next_main(ibc(k)) for k in range(7)])


def instruction_deselector_single_row(e, instruction):
    return instruction_deselector_common(e, instruction, [e.main(ibc(k)) for k in range(7)])


def initial(b):
    e = Env(b)
    clk_is_0 = e.main(C.CLK)
    ip_is_0 = e.main(C.IP)
    jsp_is_0 = e.main(C.JSP)
    jso_is_0 = e.main(C.JSO)
    jsd_is_0 = e.main(C.JSD)
    st_is_0 = [e.main(st(i)) for i in range(11)]
    op_stack_pointer_is_16 = e.main(C.OpStackPointer) - e.constant(16)

    program_digest = [e.main(st(i)) for i in range(11, 16)]
    compressed_program_digest = e.x_constant(1)
    for de in program_digest:
        compressed_program_digest = (compressed_program_digest
                                     * e.challenge(CH.CompressProgramDigestIndeterminate) + de)
    digest_ok = compressed_program_digest - e.challenge(CH.CompressedProgramDigest)

    input_init = e.aux(A.InputTableEvalArg) - e.x_constant(1)

    instruction_lookup_indeterminate = e.challenge(CH.InstructionLookupIndeterminate)
    instruction_ci_weight = e.challenge(CH.ProgramInstructionWeight)
    instruction_nia_weight = e.challenge(CH.ProgramNextInstructionWeight)
    compressed_row_for_instruction_lookup = (instruction_ci_weight * e.main(C.CI)
                                             + instruction_nia_weight * e.main(C.NIA))
    instruction_lookup_init = ((e.aux(A.InstructionLookupClientLogDerivative) - e.x_constant(0))
                               * (instruction_lookup_indeterminate - compressed_row_for_instruction_lookup)
                               - e.constant(1))

    output_init = e.aux(A.OutputTableEvalArg) - e.x_constant(1)
    op_stack_init = e.aux(A.OpStackTablePermArg) - e.x_constant(1)
    ram_init = e.aux(A.RamTablePermArg) - e.x_constant(1)

    jump_stack_indeterminate = e.challenge(CH.JumpStackIndeterminate)
    jump_stack_ci_weight = e.challenge(CH.JumpStackCiWeight)
    compressed_row_for_jump_stack_table = jump_stack_ci_weight * e.main(C.CI)
    jump_stack_init = (e.aux(A.JumpStackTablePermArg)
                       - e.x_constant(1) * (jump_stack_indeterminate - compressed_row_for_jump_stack_table))

    cjd_init = (e.aux(A.ClockJumpDifferenceLookupServerLogDerivative)
                * e.challenge(CH.ClockJumpDifferenceLookupIndeterminate)
                - e.main(C.ClockJumpDifferenceLookupMultiplicity))

    hash_selector = e.main(C.CI) - e.constant(OPCODE["hash"])
    hash_deselector = instruction_deselector_single_row(e, "hash")
    hash_input_indeterminate = e.challenge(CH.HashInputIndeterminate)
    compressed_row = e.constant(0)
    hash_input_absorbed_first_row = (e.aux(A.HashInputEvalArg)
                                     - hash_input_indeterminate * e.x_constant(1) - compressed_row)
    hash_input_default_initial = e.aux(A.HashInputEvalArg) - e.x_constant(1)
    hash_input_init = hash_selector * hash_input_default_initial + hash_deselector * hash_input_absorbed_first_row

    hash_digest_init = e.aux(A.HashDigestEvalArg) - e.x_constant(1)
    sponge_init = e.aux(A.SpongeEvalArg) - e.x_constant(1)
    u32_init = e.aux(A.U32LookupClientLogDerivative) - e.x_constant(0)

    return ([clk_is_0, ip_is_0, jsp_is_0, jso_is_0, jsd_is_0] + st_is_0
            + [digest_ok, op_stack_pointer_is_16, input_init, instruction_lookup_init, output_init, op_stack_init,
               ram_init, jump_stack_init, cjd_init, hash_input_init, hash_digest_init, sponge_init, u32_init])


def consistency(b):
    e = Env(b)
    ib_composition = (e.main(C.IB0)
                      + e.constant(1 << 1) * e.main(C.IB1)
                      + e.constant(1 << 2) * e.main(C.IB2)
                      + e.constant(1 << 3) * e.main(C.IB3)
                      + e.constant(1 << 4) * e.main(C.IB4)
                      + e.constant(1 << 5) * e.main(C.IB5)
                      + e.constant(1 << 6) * e.main(C.IB6))
    ci_corresponds_to_ib0_thru_ib6 = e.main(C.CI) - ib_composition
    ib_is_bit = [e.main(ibc(k)) * (e.main(ibc(k)) - e.constant(1)) for k in range(7)]
    is_padding_is_bit = e.main(C.IsPadding) * (e.main(C.IsPadding) - e.constant(1))
    cjd_mult_0_in_padding = (e.main(C.IsPadding) * (e.main(C.CLK) - e.constant(1))
                             * e.main(C.ClockJumpDifferenceLookupMultiplicity))
    return ib_is_bit + [is_padding_is_bit, ci_corresponds_to_ib0_thru_ib6, cjd_mult_0_in_padding]


def terminal(b):
    e = Env(b)
    return [e.main(C.CI) - e.constant(OPCODE["halt"])]


# ---- helpers (processor.rs:3075-3134) ----------------------------------------------------
def helper_variable(e, index):
    return e.cur_main(hvc(index))


def indicator_polynomial(e, index):
    one = lambda: e.constant(1)
    hv = lambda i: helper_variable(e, i)
    bits = [(index >> 3) & 1, (index >> 2) & 1, (index >> 1) & 1, index & 1]   # hv3, hv2, hv1, hv0
    acc = None
    for pos, bit in zip((3, 2, 1, 0), bits):
        f = hv(pos) if bit else (one() - hv(pos))
        acc = f if acc is None else acc * f
    return acc


# ---- instruction groups (processor.rs:436-708) -------------------------------------------
def instruction_group_decompose_arg(e):
    hv_bits = [e.cur_main(hvc(k)) * (e.cur_main(hvc(k)) - e.constant(1)) for k in range(4)]
    decomposition = (e.cur_main(C.NIA)
                     - e.constant(8) * e.cur_main(C.HV3)
                     - e.constant(4) * e.cur_main(C.HV2)
                     - e.constant(2) * e.cur_main(C.HV1)
                     - e.cur_main(C.HV0))
    return hv_bits + [decomposition]


def instruction_group_no_ram(e):
    return [e.next_aux(A.RamTablePermArg) - e.cur_aux(A.RamTablePermArg)]


def running_evaluation_for_standard_input_remains_unchanged(e):
    return e.next_aux(A.InputTableEvalArg) - e.cur_aux(A.InputTableEvalArg)


def running_evaluation_for_standard_output_remains_unchanged(e):
    return e.next_aux(A.OutputTableEvalArg) - e.cur_aux(A.OutputTableEvalArg)


def instruction_group_no_io(e):
    return [running_evaluation_for_standard_input_remains_unchanged(e),
            running_evaluation_for_standard_output_remains_unchanged(e)]


def instruction_group_keep_op_stack_height(e):
    op_stack_pointer_curr = e.cur_main(C.OpStackPointer)
    op_stack_pointer_next = e.next_main(C.OpStackPointer)
    osp_remains_unchanged = op_stack_pointer_next - op_stack_pointer_curr
    perm_arg_curr = e.cur_aux(A.OpStackTablePermArg)
    perm_arg_next = e.next_aux(A.OpStackTablePermArg)
    perm_arg_remains_unchanged = perm_arg_next - perm_arg_curr
    return [osp_remains_unchanged, perm_arg_remains_unchanged]


def instruction_group_op_stack_remains_except_top_n(e, n):
    assert n <= NUM_OP_STACK_REGISTERS
    next_stack = [e.next_main(st(i)) for i in range(16)]
    curr_stack = [e.cur_main(st(i)) for i in range(16)]

    def compress_stack_except_top_n(stack):
        return msum(stack_weight(e, i) * s for i, s in list(enumerate(stack))[n:])

    all_but_n_top_elements_remain = compress_stack_except_top_n(next_stack) - compress_stack_except_top_n(curr_stack)
    constraints = instruction_group_keep_op_stack_height(e)
    constraints.append(all_but_n_top_elements_remain)
    return constraints


def instruction_group_keep_op_stack(e):
    return instruction_group_op_stack_remains_except_top_n(e, 0)


def single_factor_for_permutation_argument_with_op_stack_table(e, shorter_is_next, op_stack_pointer_offset):
    row_with_shorter_stack = e.next_main if shorter_is_next else e.cur_main
    max_stack_element_index = 15
    stack_element_index = max_stack_element_index - op_stack_pointer_offset
    underflow_element = row_with_shorter_stack(st(stack_element_index))
    op_stack_pointer = row_with_shorter_stack(C.OpStackPointer)
    offset = e.constant(op_stack_pointer_offset)
    offset_op_stack_pointer = op_stack_pointer + offset
    compressed_row = (e.challenge(CH.OpStackClkWeight) * e.cur_main(C.CLK)
                      + e.challenge(CH.OpStackIb1Weight) * e.cur_main(C.IB1)
                      + e.challenge(CH.OpStackPointerWeight) * offset_op_stack_pointer
                      + e.challenge(CH.OpStackFirstUnderflowElementWeight) * underflow_element)
    return e.challenge(CH.OpStackIndeterminate) - compressed_row


def running_product_op_stack_accounts_for_growing_stack_by(e, n):
    factor = e.constant(1)
    for off in range(n):
        factor = factor * single_factor_for_permutation_argument_with_op_stack_table(e, False, off)
    return e.next_aux(A.OpStackTablePermArg) - e.cur_aux(A.OpStackTablePermArg) * factor


def running_product_op_stack_accounts_for_shrinking_stack_by(e, n):
    factor = e.constant(1)
    for off in range(n):
        factor = factor * single_factor_for_permutation_argument_with_op_stack_table(e, True, off)
    return e.next_aux(A.OpStackTablePermArg) - e.cur_aux(A.OpStackTablePermArg) * factor


def instruction_group_grow_op_stack_and_top_two_elements_unconstrained(e):
    out = [e.next_main(st(i + 1)) - e.cur_main(st(i)) for i in range(1, 15)]
    out.append(e.next_main(C.OpStackPointer) - e.cur_main(C.OpStackPointer) - e.constant(1))
    out.append(running_product_op_stack_accounts_for_growing_stack_by(e, 1))
    return out


def instruction_group_grow_op_stack(e):
    specific = [e.next_main(C.ST1) - e.cur_main(C.ST0)]
    inherited = instruction_group_grow_op_stack_and_top_two_elements_unconstrained(e)
    return specific + inherited


def instruction_group_op_stack_shrinks_and_top_three_elements_unconstrained(e):
    out = [e.next_main(st(i)) - e.cur_main(st(i + 1)) for i in range(3, 15)]
    out.append(e.next_main(C.OpStackPointer) - e.cur_main(C.OpStackPointer) + e.constant(1))
    out.append(running_product_op_stack_accounts_for_shrinking_stack_by(e, 1))
    return out


def instruction_group_binop(e):
    specific = [e.next_main(C.ST1) - e.cur_main(C.ST2), e.next_main(C.ST2) - e.cur_main(C.ST3)]
    inherited = instruction_group_op_stack_shrinks_and_top_three_elements_unconstrained(e)
    return specific + inherited


def instruction_group_shrink_op_stack(e):
    specific = [e.next_main(C.ST0) - e.cur_main(C.ST1)]
    inherited = instruction_group_binop(e)
    return specific + inherited


def instruction_group_keep_jump_stack(e):
    jsp = e.next_main(C.JSP) - e.cur_main(C.JSP)
    jso = e.next_main(C.JSO) - e.cur_main(C.JSO)
    jsd = e.next_main(C.JSD) - e.cur_main(C.JSD)
    return [jsp, jso, jsd]


def instruction_group_step_1(e):
    ip_inc = e.next_main(C.IP) - e.cur_main(C.IP) - e.constant(1)
    return instruction_group_keep_jump_stack(e) + [ip_inc]


def instruction_group_step_2(e):
    ip_inc = e.next_main(C.IP) - e.cur_main(C.IP) - e.constant(2)
    return instruction_group_keep_jump_stack(e) + [ip_inc]


# ---- stack growth / shrink helpers (processor.rs:2354-2478) ---------------------------------
def combine_mutually_exclusive_constraint_groups(e, groups):
    num_constraints = max((len(g) for g in groups), default=0)
    combined = []
    for i in range(num_constraints):
        acc = e.constant(0)
        for g in groups:
            if i < len(g):
                acc = acc + g[i]
        combined.append(acc)
    return combined


def constraints_for_shrinking_stack_by(e, n):
    new_stack = [e.next_main(st(i)) for i in range(16 - n)]            # dropping_back(n) before map
    old_stack_with_top_n_removed = [e.cur_main(st(i)) for i in range(n, 16)]

    def compress(stack):
        assert len(stack) == 16 - n
        return msum(stack_weight(e, i) * s for i, s in enumerate(stack))

    compressed_new_stack = compress(new_stack)
    compressed_old_stack = compress(old_stack_with_top_n_removed)
    op_stack_pointer_shrinks_by_n = e.next_main(C.OpStackPointer) - e.cur_main(C.OpStackPointer) + e.constant(n)
    new_is_old_with_top_n_removed = compressed_new_stack - compressed_old_stack
    return [op_stack_pointer_shrinks_by_n, new_is_old_with_top_n_removed,
            running_product_op_stack_accounts_for_shrinking_stack_by(e, n)]


def constraints_for_growing_stack_by(e, n):
    new_stack = [e.next_main(st(i)) for i in range(n, 16)]
    # stack().map(curr_row).dropping_back(n): the mapped closure runs for the dropped elements
    # first, from the back (itertools::dropping_back), then for the kept ones in order.
    for i in range(15, 15 - n, -1):
        e.cur_main(st(i))
    old_stack_with_top_n_added = [e.cur_main(st(i)) for i in range(16 - n)]

    def compress(stack):
        assert len(stack) == 16 - n
        return msum(stack_weight(e, i) * s for i, s in enumerate(stack))

    compressed_new_stack = compress(new_stack)
    compressed_old_stack = compress(old_stack_with_top_n_added)
    op_stack_pointer_grows_by_n = e.next_main(C.OpStackPointer) - e.cur_main(C.OpStackPointer) - e.constant(n)
    new_is_old_with_top_n_added = compressed_new_stack - compressed_old_stack
    return [op_stack_pointer_grows_by_n, new_is_old_with_top_n_added,
            running_product_op_stack_accounts_for_growing_stack_by(e, n)]


def conditional_constraints_for_shrinking_stack_by(e, n):
    return [indicator_polynomial(e, n) * c for c in constraints_for_shrinking_stack_by(e, n)]


def conditional_constraints_for_growing_stack_by(e, n):
    return [indicator_polynomial(e, n) * c for c in constraints_for_growing_stack_by(e, n)]


def stack_shrinks_by_any_of(e, shrinkages):
    groups = [conditional_constraints_for_shrinking_stack_by(e, n) for n in shrinkages]
    return combine_mutually_exclusive_constraint_groups(e, groups)


def stack_grows_by_any_of(e, growths):
    groups = [conditional_constraints_for_growing_stack_by(e, n) for n in growths]
    return combine_mutually_exclusive_constraint_groups(e, groups)


def prohibit_any_illegal_number_of_words(e):
    # illegal_values().map(indicator) is an eager array map; then summed
    polys = [indicator_polynomial(e, n) for n in ILLEGAL_NUMBER_OF_WORDS]
    return [msum(polys)]


def constraints_for_shrinking_stack_by_3_and_top_3_unconstrained(e):
    out = [e.next_main(st(i)) - e.cur_main(st(i + 3)) for i in range(3, 13)]
    out.append(e.next_main(C.OpStackPointer) - e.cur_main(C.OpStackPointer) + e.constant(3))
    out.append(running_product_op_stack_accounts_for_shrinking_stack_by(e, 3))
    return out


# ---- RAM helpers (processor.rs:1923-1950, 2494-2650) ---------------------------------------
def read_from_ram_to(e, ram_pointers, destinations):
    def compress_row(ram_pointer, destination):
        return (e.cur_main(C.CLK) * e.challenge(CH.RamClkWeight)
                + e.constant(RAM_INSTRUCTION_TYPE_READ) * e.challenge(CH.RamInstructionTypeWeight)
                + ram_pointer * e.challenge(CH.RamPointerWeight)
                + destination * e.challenge(CH.RamValueWeight))

    factor = None
    for rp, d in zip(ram_pointers, destinations):
        compressed_row = compress_row(rp, d)
        f = e.challenge(CH.RamIndeterminate) - compressed_row
        factor = f if factor is None else factor * f
    if factor is None:
        factor = e.constant(1)
    return e.cur_aux(A.RamTablePermArg) * factor - e.next_aux(A.RamTablePermArg)


def single_factor_for_permutation_argument_with_ram_table(e, longer_is_next, instruction_type, ram_pointer_offset):
    row_with_longer_stack = e.next_main if longer_is_next else e.cur_main
    num_ram_pointers = 1
    ram_value_index = ram_pointer_offset + num_ram_pointers
    ram_value = row_with_longer_stack(st(ram_value_index))
    additional_offset = 1 if instruction_type == RAM_INSTRUCTION_TYPE_READ else 0
    ram_pointer = row_with_longer_stack(C.ST0)
    offset = e.constant(additional_offset + ram_pointer_offset)
    offset_ram_pointer = ram_pointer + offset
    compressed_row = (e.cur_main(C.CLK) * e.challenge(CH.RamClkWeight)
                      + e.constant(instruction_type) * e.challenge(CH.RamInstructionTypeWeight)
                      + offset_ram_pointer * e.challenge(CH.RamPointerWeight)
                      + ram_value * e.challenge(CH.RamValueWeight))
    return e.challenge(CH.RamIndeterminate) - compressed_row


def running_product_ram_accounts_for_writing_n_elements(e, n):
    factor = e.constant(1)
    for off in range(n):
        factor = factor * single_factor_for_permutation_argument_with_ram_table(e, False, RAM_INSTRUCTION_TYPE_WRITE, off)
    return e.next_aux(A.RamTablePermArg) - e.cur_aux(A.RamTablePermArg) * factor


def running_product_ram_accounts_for_reading_n_elements(e, n):
    factor = e.constant(1)
    for off in range(n):
        factor = factor * single_factor_for_permutation_argument_with_ram_table(e, True, RAM_INSTRUCTION_TYPE_READ, off)
    return e.next_aux(A.RamTablePermArg) - e.cur_aux(A.RamTablePermArg) * factor


def shrink_stack_by_n_and_write_n_elements_to_ram(e, n):
    osp = e.next_main(C.OpStackPointer) - e.cur_main(C.OpStackPointer) + e.constant(n)
    ram_pointer_grows_by_n = e.next_main(C.ST0) - e.cur_main(C.ST0) - e.constant(n)
    constraints = [osp, ram_pointer_grows_by_n,
                   running_product_op_stack_accounts_for_shrinking_stack_by(e, n),
                   running_product_ram_accounts_for_writing_n_elements(e, n)]
    for i in range(n + 1, 16):
        constraints.append(e.next_main(st(i - n)) - e.cur_main(st(i)))
    return constraints


def grow_stack_by_n_and_read_n_elements_from_ram(e, n):
    osp = e.next_main(C.OpStackPointer) - e.cur_main(C.OpStackPointer) - e.constant(n)
    ram_pointer_shrinks_by_n = e.next_main(C.ST0) - e.cur_main(C.ST0) + e.constant(n)
    constraints = [osp, ram_pointer_shrinks_by_n,
                   running_product_op_stack_accounts_for_growing_stack_by(e, n),
                   running_product_ram_accounts_for_reading_n_elements(e, n)]
    for i in range(1, 16 - n):
        constraints.append(e.next_main(st(i + n)) - e.cur_main(st(i)))
    return constraints


def conditional_constraints_for_writing_n_elements_to_ram(e, n):
    return [indicator_polynomial(e, n) * c for c in shrink_stack_by_n_and_write_n_elements_to_ram(e, n)]


def conditional_constraints_for_reading_n_elements_from_ram(e, n):
    return [indicator_polynomial(e, n) * c for c in grow_stack_by_n_and_read_n_elements_from_ram(e, n)]


def write_to_ram_any_of(e, ns):
    return combine_mutually_exclusive_constraint_groups(
        e, [conditional_constraints_for_writing_n_elements_to_ram(e, n) for n in ns])


def read_from_ram_any_of(e, ns):
    return combine_mutually_exclusive_constraint_groups(
        e, [conditional_constraints_for_reading_n_elements_from_ram(e, n) for n in ns])


# ---- IO helpers (processor.rs:2242-2290) -----------------------------------------------------
def grow_stack_by_n_and_read_n_symbols_from_input(e, n):
    indeterminate = lambda: e.challenge(CH.StandardInputIndeterminate)
    running_evaluation = e.cur_aux(A.InputTableEvalArg)
    for i in reversed(range(n)):
        running_evaluation = indeterminate() * running_evaluation + e.next_main(st(i))
    running_evaluation_update = e.next_aux(A.InputTableEvalArg) - running_evaluation
    conditional = indicator_polynomial(e, n) * running_evaluation_update
    constraints = conditional_constraints_for_growing_stack_by(e, n)
    constraints.append(conditional)
    return constraints


def shrink_stack_by_n_and_write_n_symbols_to_output(e, n):
    indeterminate = lambda: e.challenge(CH.StandardOutputIndeterminate)
    running_evaluation = e.cur_aux(A.OutputTableEvalArg)
    for i in range(n):
        running_evaluation = indeterminate() * running_evaluation + e.cur_main(st(i))
    running_evaluation_update = e.next_aux(A.OutputTableEvalArg) - running_evaluation
    conditional = indicator_polynomial(e, n) * running_evaluation_update
    constraints = conditional_constraints_for_shrinking_stack_by(e, n)
    constraints.append(conditional)
    return constraints


# ---- X-field products (processor.rs:1952-1978) -----------------------------------------------
def xx_product(x, y):
    x_0, x_1, x_2 = x
    y_0, y_1, y_2 = y
    z0 = x_0 * y_0
    z1 = x_1 * y_0 + x_0 * y_1
    z2 = x_2 * y_0 + x_1 * y_1 + x_0 * y_2
    z3 = x_2 * y_1 + x_1 * y_2
    z4 = x_2 * y_2
    return [z0 - z3, z1 - z4 + z3, z2 + z4]


def xb_product(x, y):
    x_0, x_1, x_2 = x
    return [x_0 * y, x_1 * y, x_2 * y]


# ---- per-instruction constraints (processor.rs:820-2102) -------------------------------------
def instruction_pop(e):
    return (instruction_group_step_2(e) + instruction_group_decompose_arg(e)
            + stack_shrinks_by_any_of(e, LEGAL_NUMBER_OF_WORDS) + prohibit_any_illegal_number_of_words(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_push(e):
    specific = [e.next_main(C.ST0) - e.cur_main(C.NIA)]
    return (specific + instruction_group_grow_op_stack(e) + instruction_group_step_2(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_divine(e):
    return (instruction_group_step_2(e) + instruction_group_decompose_arg(e)
            + stack_grows_by_any_of(e, LEGAL_NUMBER_OF_WORDS) + prohibit_any_illegal_number_of_words(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def _stack_permutation_instruction(e, permuted):
    """shared shape of pick / place / swap (processor.rs:867-1029)"""
    stack = list(range(16))
    next_stack = [e.next_main(st(i)) for i in stack]

    def compress(stk):
        assert len(stk) == 16
        return msum(stack_weight(e, i) * s for i, s in enumerate(stk))

    def term(i):
        return (indicator_polynomial(e, i)
                * (compress(next_stack) - compress([e.cur_main(st(j)) for j in permuted(stack, i)])))

    total = msum(term(i) for i in range(16))
    return ([total] + instruction_group_decompose_arg(e) + instruction_group_step_2(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e) + instruction_group_keep_op_stack_height(e))


def instruction_pick(e):
    def permuted(stack, i):
        s = list(stack); new_top = s.pop(i); s.insert(0, new_top); return s
    return _stack_permutation_instruction(e, permuted)


def instruction_place(e):
    def permuted(stack, i):
        s = list(stack); old_top = s.pop(0); s.insert(i, old_top); return s
    return _stack_permutation_instruction(e, permuted)


def instruction_swap(e):
    def permuted(stack, i):
        s = list(stack); s[0], s[i] = s[i], s[0]; return s
    return _stack_permutation_instruction(e, permuted)


def instruction_dup(e):
    duplicate_element = lambda i: indicator_polynomial(e, i) * (e.next_main(C.ST0) - e.cur_main(st(i)))
    duplicate_indicated_element = msum(duplicate_element(i) for i in range(16))
    return ([duplicate_indicated_element] + instruction_group_decompose_arg(e) + instruction_group_step_2(e)
            + instruction_group_grow_op_stack(e) + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_halt(e):
    specific = [e.next_main(C.CI) - e.cur_main(C.CI)]
    return (specific + instruction_group_step_1(e) + instruction_group_keep_op_stack(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_nop(e):
    return (instruction_group_step_1(e) + instruction_group_keep_op_stack(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def next_instruction_range_check_constraints_for_instruction_skiz(e):
    is_0_or_1 = lambda v: e.cur_main(v) * (e.cur_main(v) - e.constant(1))
    is_0_to_3 = lambda v: (e.cur_main(v) * (e.cur_main(v) - e.constant(1)) * (e.cur_main(v) - e.constant(2))
                           * (e.cur_main(v) - e.constant(3)))
    return [is_0_or_1(C.HV1), is_0_to_3(C.HV2), is_0_to_3(C.HV3), is_0_to_3(C.HV4), is_0_to_3(C.HV5)]


def instruction_skiz(e):
    one = lambda: e.constant(1)
    hv0_is_inverse_of_st0 = e.cur_main(C.HV0) * e.cur_main(C.ST0) - one()
    c0 = hv0_is_inverse_of_st0 * e.cur_main(C.HV0)
    c1 = hv0_is_inverse_of_st0 * e.cur_main(C.ST0)
    nia_decomposes_to_hvs = (e.cur_main(C.NIA) - e.cur_main(C.HV1)
                             - e.constant(1 << 1) * e.cur_main(C.HV2)
                             - e.constant(1 << 3) * e.cur_main(C.HV3)
                             - e.constant(1 << 5) * e.cur_main(C.HV4)
                             - e.constant(1 << 7) * e.cur_main(C.HV5))
    ip_case_1 = (e.next_main(C.IP) - e.cur_main(C.IP) - e.constant(1)) * e.cur_main(C.ST0)
    ip_case_2 = ((e.next_main(C.IP) - e.cur_main(C.IP) - e.constant(2))
                 * (e.cur_main(C.ST0) * e.cur_main(C.HV0) - one()) * (e.cur_main(C.HV1) - one()))
    ip_case_3 = ((e.next_main(C.IP) - e.cur_main(C.IP) - e.constant(3))
                 * (e.cur_main(C.ST0) * e.cur_main(C.HV0) - one()) * e.cur_main(C.HV1))
    ip_incr = ip_case_1 + ip_case_2 + ip_case_3
    specific = [c0, c1, nia_decomposes_to_hvs, ip_incr]
    return (specific + next_instruction_range_check_constraints_for_instruction_skiz(e)
            + instruction_group_keep_jump_stack(e) + instruction_group_shrink_op_stack(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_call(e):
    jsp_incr_1 = e.next_main(C.JSP) - e.cur_main(C.JSP) - e.constant(1)
    jso_becomes_ip_plus_2 = e.next_main(C.JSO) - e.cur_main(C.IP) - e.constant(2)
    jsd_becomes_nia = e.next_main(C.JSD) - e.cur_main(C.NIA)
    ip_becomes_nia = e.next_main(C.IP) - e.cur_main(C.NIA)
    specific = [jsp_incr_1, jso_becomes_ip_plus_2, jsd_becomes_nia, ip_becomes_nia]
    return specific + instruction_group_keep_op_stack(e) + instruction_group_no_ram(e) + instruction_group_no_io(e)


def instruction_return(e):
    jsp_decrements_by_1 = e.next_main(C.JSP) - e.cur_main(C.JSP) + e.constant(1)
    ip_is_set_to_jso = e.next_main(C.IP) - e.cur_main(C.JSO)
    specific = [jsp_decrements_by_1, ip_is_set_to_jso]
    return specific + instruction_group_keep_op_stack(e) + instruction_group_no_ram(e) + instruction_group_no_io(e)


def instruction_recurse(e):
    specific = [e.next_main(C.IP) - e.cur_main(C.JSD)]
    return (specific + instruction_group_keep_jump_stack(e) + instruction_group_keep_op_stack(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_recurse_or_return(e):
    one = lambda: e.constant(1)
    st5_eq_st6 = lambda: e.cur_main(C.HV0) * (e.cur_main(C.ST6) - e.cur_main(C.ST5))
    st5_neq_st6 = lambda: one() - st5_eq_st6()
    c0 = st5_neq_st6() * e.cur_main(C.HV0)
    c1 = st5_neq_st6() * (e.cur_main(C.ST6) - e.cur_main(C.ST5))
    specific = [c0, c1]
    maybe_return = [st5_neq_st6() * (e.next_main(C.IP) - e.cur_main(C.JSO)),
                    st5_neq_st6() * (e.next_main(C.JSP) - e.cur_main(C.JSP) + one())]
    maybe_recurse = [st5_eq_st6() * (e.next_main(C.IP) - e.cur_main(C.JSD)),
                     st5_eq_st6() * (e.next_main(C.JSP) - e.cur_main(C.JSP)),
                     st5_eq_st6() * (e.next_main(C.JSO) - e.cur_main(C.JSO)),
                     st5_eq_st6() * (e.next_main(C.JSD) - e.cur_main(C.JSD))]
    specific += combine_mutually_exclusive_constraint_groups(e, [maybe_return, maybe_recurse])
    return specific + instruction_group_keep_op_stack(e) + instruction_group_no_ram(e) + instruction_group_no_io(e)


def instruction_assert(e):
    specific = [e.cur_main(C.ST0) - e.constant(1)]
    return (specific + instruction_group_step_1(e) + instruction_group_shrink_op_stack(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_read_mem(e):
    return (instruction_group_step_2(e) + instruction_group_decompose_arg(e)
            + read_from_ram_any_of(e, LEGAL_NUMBER_OF_WORDS) + prohibit_any_illegal_number_of_words(e)
            + instruction_group_no_io(e))


def instruction_write_mem(e):
    return (instruction_group_step_2(e) + instruction_group_decompose_arg(e)
            + write_to_ram_any_of(e, LEGAL_NUMBER_OF_WORDS) + prohibit_any_illegal_number_of_words(e)
            + instruction_group_no_io(e))


def instruction_hash(e):
    shrinks = [e.next_main(st(i)) - e.cur_main(st(i + 5)) for i in range(5, 11)]
    shrinks.append(e.next_main(C.OpStackPointer) - e.cur_main(C.OpStackPointer) + e.constant(5))
    shrinks.append(running_product_op_stack_accounts_for_shrinking_stack_by(e, 5))
    return instruction_group_step_1(e) + shrinks + instruction_group_no_ram(e) + instruction_group_no_io(e)


def instruction_merkle_step_shared_constraints(e):
    one = lambda: e.constant(1)
    hv5_is_0_or_1 = e.cur_main(C.HV5) * (e.cur_main(C.HV5) - one())
    new_st5 = e.constant(2) * e.next_main(C.ST5) + e.cur_main(C.HV5) - e.cur_main(C.ST5)
    return [hv5_is_0_or_1, new_st5] + instruction_group_step_1(e) + instruction_group_no_io(e)


def instruction_merkle_step(e):
    return (instruction_merkle_step_shared_constraints(e) + instruction_group_op_stack_remains_except_top_n(e, 6)
            + instruction_group_no_ram(e))


def instruction_merkle_step_mem(e):
    ram_pointers = [e.cur_main(C.ST7) + e.constant(i) for i in range(5)]
    ram_read_destinations = [e.cur_main(hvc(i)) for i in range(5)]
    read_from_ram_to_hvs = read_from_ram_to(e, ram_pointers, ram_read_destinations)
    st6_does_not_change = e.next_main(C.ST6) - e.cur_main(C.ST6)
    st7_increments_by_5 = e.next_main(C.ST7) - e.cur_main(C.ST7) - e.constant(5)
    st6_and_st7 = stack_weight(e, 6) * st6_does_not_change + stack_weight(e, 7) * st7_increments_by_5
    return ([st6_and_st7, read_from_ram_to_hvs] + instruction_merkle_step_shared_constraints(e)
            + instruction_group_op_stack_remains_except_top_n(e, 8))


def instruction_assert_vector(e):
    specific = [e.cur_main(st(i + 5)) - e.cur_main(st(i)) for i in range(5)]
    return (specific + instruction_group_step_1(e) + constraints_for_shrinking_stack_by(e, 5)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_sponge_init(e):
    return (instruction_group_step_1(e) + instruction_group_keep_op_stack(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_sponge_absorb(e):
    return (instruction_group_step_1(e) + constraints_for_shrinking_stack_by(e, 10)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_sponge_absorb_mem(e):
    increment_ram_pointer = e.next_main(C.ST0) - e.cur_main(C.ST0) - e.constant(RATE)
    ram_pointers = [e.cur_main(C.ST0) + e.constant(i) for i in range(10)]
    ram_read_destinations = ([e.next_main(st(i)) for i in range(1, 5)] + [e.cur_main(hvc(i)) for i in range(6)])
    read_from_ram = read_from_ram_to(e, ram_pointers, ram_read_destinations)
    return ([increment_ram_pointer, read_from_ram] + instruction_group_step_1(e)
            + instruction_group_op_stack_remains_except_top_n(e, 5) + instruction_group_no_io(e))


def instruction_sponge_squeeze(e):
    return (instruction_group_step_1(e) + constraints_for_growing_stack_by(e, 10)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_add(e):
    specific = [e.next_main(C.ST0) - e.cur_main(C.ST0) - e.cur_main(C.ST1)]
    return (specific + instruction_group_step_1(e) + instruction_group_binop(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_addi(e):
    specific = [e.next_main(C.ST0) - e.cur_main(C.ST0) - e.cur_main(C.NIA)]
    return (specific + instruction_group_step_2(e) + instruction_group_op_stack_remains_except_top_n(e, 1)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_mul(e):
    specific = [e.next_main(C.ST0) - e.cur_main(C.ST0) * e.cur_main(C.ST1)]
    return (specific + instruction_group_step_1(e) + instruction_group_binop(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_invert(e):
    specific = [e.next_main(C.ST0) * e.cur_main(C.ST0) - e.constant(1)]
    return (specific + instruction_group_step_1(e) + instruction_group_op_stack_remains_except_top_n(e, 1)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_eq(e):
    one = lambda: e.constant(1)
    st0_eq_st1 = lambda: one() - e.cur_main(C.HV0) * (e.cur_main(C.ST1) - e.cur_main(C.ST0))
    c0 = e.cur_main(C.HV0) * st0_eq_st1()
    c1 = (e.cur_main(C.ST1) - e.cur_main(C.ST0)) * st0_eq_st1()
    c2 = e.next_main(C.ST0) - st0_eq_st1()
    return ([c0, c1, c2] + instruction_group_step_1(e) + instruction_group_binop(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_split(e):
    one = lambda: e.constant(1)
    c0 = e.cur_main(C.ST0) - (e.constant(1 << 32) * e.next_main(C.ST1) + e.next_main(C.ST0))
    hv0 = e.cur_main(C.HV0)
    hi = e.next_main(C.ST1)
    lo = e.next_main(C.ST0)
    ffff_ffff = e.constant(0xFFFF_FFFF)
    c1 = lo * (hv0 * (hi - ffff_ffff) - one())
    return ([c0, c1] + instruction_group_grow_op_stack_and_top_two_elements_unconstrained(e)
            + instruction_group_step_1(e) + instruction_group_no_ram(e) + instruction_group_no_io(e))


def _step1_binop(e):
    return (instruction_group_step_1(e) + instruction_group_binop(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def _step1_unop(e):
    return (instruction_group_step_1(e) + instruction_group_op_stack_remains_except_top_n(e, 1)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


instruction_lt = instruction_and = instruction_xor = instruction_pow = _step1_binop
instruction_log_2_floor = instruction_pop_count = _step1_unop


def instruction_div_mod(e):
    specific = [e.cur_main(C.ST0) - e.cur_main(C.ST1) * e.next_main(C.ST1) - e.next_main(C.ST0)]
    return (specific + instruction_group_step_1(e) + instruction_group_op_stack_remains_except_top_n(e, 2)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_xx_add(e):
    c0 = e.next_main(C.ST0) - e.cur_main(C.ST0) - e.cur_main(C.ST3)
    c1 = e.next_main(C.ST1) - e.cur_main(C.ST1) - e.cur_main(C.ST4)
    c2 = e.next_main(C.ST2) - e.cur_main(C.ST2) - e.cur_main(C.ST5)
    return ([c0, c1, c2] + constraints_for_shrinking_stack_by_3_and_top_3_unconstrained(e)
            + instruction_group_step_1(e) + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_xx_mul(e):
    x0, x1, x2, y0, y1, y2 = [e.cur_main(st(i)) for i in range(6)]
    c0, c1, c2 = xx_product([x0, x1, x2], [y0, y1, y2])
    specific = [e.next_main(C.ST0) - c0, e.next_main(C.ST1) - c1, e.next_main(C.ST2) - c2]
    return (specific + constraints_for_shrinking_stack_by_3_and_top_3_unconstrained(e)
            + instruction_group_step_1(e) + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_xinv(e):
    c0 = (e.cur_main(C.ST0) * e.next_main(C.ST0)
          - e.cur_main(C.ST2) * e.next_main(C.ST1)
          - e.cur_main(C.ST1) * e.next_main(C.ST2)
          - e.constant(1))
    c1 = (e.cur_main(C.ST1) * e.next_main(C.ST0)
          + e.cur_main(C.ST0) * e.next_main(C.ST1)
          - e.cur_main(C.ST2) * e.next_main(C.ST2)
          + e.cur_main(C.ST2) * e.next_main(C.ST1)
          + e.cur_main(C.ST1) * e.next_main(C.ST2))
    c2 = (e.cur_main(C.ST2) * e.next_main(C.ST0)
          + e.cur_main(C.ST1) * e.next_main(C.ST1)
          + e.cur_main(C.ST0) * e.next_main(C.ST2)
          + e.cur_main(C.ST2) * e.next_main(C.ST2))
    return ([c0, c1, c2] + instruction_group_op_stack_remains_except_top_n(e, 3)
            + instruction_group_step_1(e) + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_xb_mul(e):
    x, y0, y1, y2 = [e.cur_main(st(i)) for i in range(4)]
    c0, c1, c2 = xb_product([y0, y1, y2], x)
    specific = [e.next_main(C.ST0) - c0, e.next_main(C.ST1) - c1, e.next_main(C.ST2) - c2]
    return (specific + instruction_group_op_stack_shrinks_and_top_three_elements_unconstrained(e)
            + instruction_group_step_1(e) + instruction_group_no_ram(e) + instruction_group_no_io(e))


def instruction_read_io(e):
    groups = [grow_stack_by_n_and_read_n_symbols_from_input(e, n) for n in LEGAL_NUMBER_OF_WORDS]
    read_any = combine_mutually_exclusive_constraint_groups(e, groups)
    return (instruction_group_step_2(e) + instruction_group_decompose_arg(e) + read_any
            + prohibit_any_illegal_number_of_words(e) + instruction_group_no_ram(e)
            + [running_evaluation_for_standard_output_remains_unchanged(e)])


def instruction_write_io(e):
    groups = [shrink_stack_by_n_and_write_n_symbols_to_output(e, n) for n in LEGAL_NUMBER_OF_WORDS]
    write_any = combine_mutually_exclusive_constraint_groups(e, groups)
    return (instruction_group_step_2(e) + instruction_group_decompose_arg(e) + write_any
            + prohibit_any_illegal_number_of_words(e) + instruction_group_no_ram(e)
            + [running_evaluation_for_standard_input_remains_unchanged(e)])


def _horner_step(e, extension):
    if not extension:
        read_from_ram = read_from_ram_to(e, [e.cur_main(C.ST5)], [e.cur_main(C.HV0)])
    else:
        ram_pointers = [e.cur_main(C.ST5) - e.constant(i) for i in range(3)]
        read_from_ram = read_from_ram_to(e, ram_pointers, [e.cur_main(C.HV2), e.cur_main(C.HV1), e.cur_main(C.HV0)])
    indeterminate = [e.cur_main(C.ST0), e.cur_main(C.ST1), e.cur_main(C.ST2)]
    evaluation = [e.cur_main(C.ST7), e.cur_main(C.ST8), e.cur_main(C.ST9)]
    product_0, product_1, product_2 = xx_product(indeterminate, evaluation)
    sw = lambda i: stack_weight(e, i)
    if not extension:
        curr_stack_compressed = (sw(0) * e.cur_main(C.ST0) + sw(1) * e.cur_main(C.ST1) + sw(2) * e.cur_main(C.ST2)
                                 + sw(3) * e.cur_main(C.ST3) + sw(4) * e.cur_main(C.ST4)
                                 + sw(5) * (e.cur_main(C.ST5) - e.constant(1))
                                 + sw(6) * e.cur_main(C.ST6)
                                 + sw(7) * (product_0 + e.cur_main(C.HV0))
                                 + sw(8) * product_1
                                 + sw(9) * product_2)
    else:
        curr_stack_compressed = (sw(0) * e.cur_main(C.ST0) + sw(1) * e.cur_main(C.ST1) + sw(2) * e.cur_main(C.ST2)
                                 + sw(3) * e.cur_main(C.ST3) + sw(4) * e.cur_main(C.ST4)
                                 + sw(5) * (e.cur_main(C.ST5) - e.constant(3))
                                 + sw(6) * e.cur_main(C.ST6)
                                 + sw(7) * (product_0 + e.cur_main(C.HV0))
                                 + sw(8) * (product_1 + e.cur_main(C.HV1))
                                 + sw(9) * (product_2 + e.cur_main(C.HV2)))
    next_stack_compressed = msum(sw(i) * e.next_main(st(i)) for i in range(10))
    stack_changes_correctly = next_stack_compressed - curr_stack_compressed
    return ([stack_changes_correctly, read_from_ram] + instruction_group_no_io(e) + instruction_group_step_1(e)
            + instruction_group_op_stack_remains_except_top_n(e, 10))


def instruction_b_horner_step(e): return _horner_step(e, False)
def instruction_x_horner_step(e): return _horner_step(e, True)


INSTRUCTION_CONSTRAINTS = {
    "pop": instruction_pop, "push": instruction_push, "divine": instruction_divine, "pick": instruction_pick,
    "place": instruction_place, "dup": instruction_dup, "swap": instruction_swap, "halt": instruction_halt,
    "nop": instruction_nop, "skiz": instruction_skiz, "call": instruction_call, "return": instruction_return,
    "recurse": instruction_recurse, "recurse_or_return": instruction_recurse_or_return, "assert": instruction_assert,
    "read_mem": instruction_read_mem, "write_mem": instruction_write_mem, "hash": instruction_hash,
    "assert_vector": instruction_assert_vector, "sponge_init": instruction_sponge_init,
    "sponge_absorb": instruction_sponge_absorb, "sponge_absorb_mem": instruction_sponge_absorb_mem,
    "sponge_squeeze": instruction_sponge_squeeze, "add": instruction_add, "addi": instruction_addi,
    "mul": instruction_mul, "invert": instruction_invert, "eq": instruction_eq, "split": instruction_split,
    "lt": instruction_lt, "and": instruction_and, "xor": instruction_xor, "log_2_floor": instruction_log_2_floor,
    "pow": instruction_pow, "div_mod": instruction_div_mod, "pop_count": instruction_pop_count,
    "xx_add": instruction_xx_add, "xx_mul": instruction_xx_mul, "x_invert": instruction_xinv,
    "xb_mul": instruction_xb_mul, "read_io": instruction_read_io, "write_io": instruction_write_io,
    "merkle_step": instruction_merkle_step, "merkle_step_mem": instruction_merkle_step_mem,
    "b_horner_step": instruction_b_horner_step, "x_horner_step": instruction_x_horner_step,
}


def transition_constraints_for_instruction(e, instruction):
    return INSTRUCTION_CONSTRAINTS[instruction](e)


# ---- combination (processor.rs:363-434) -------------------------------------------------------
def combine_instruction_constraints_with_deselectors(e, deselectors, all_tc):
    max_number_of_constraints = max(len(tc) for tc in all_tc)
    zero_poly = e.constant(0)
    transposed = [[tc[idx] if idx < len(tc) else zero_poly for tc in all_tc] for idx in range(max_number_of_constraints)]
    return [msum(d * tc for d, tc in zip(deselectors, row)) for row in transposed]


def padding_row_constraints(e):
    return ([e.next_main(C.IP) - e.cur_main(C.IP), e.next_main(C.CI) - e.cur_main(C.CI),
             e.next_main(C.NIA) - e.cur_main(C.NIA)]
            + instruction_group_keep_jump_stack(e) + instruction_group_keep_op_stack(e)
            + instruction_group_no_ram(e) + instruction_group_no_io(e))


def combine_transition_constraints_with_padding_constraints(e, instruction_transition_constraints):
    padding = padding_row_constraints(e)
    padding_row_deselector = e.constant(1) - e.next_main(C.IsPadding)
    padding_row_selector = e.next_main(C.IsPadding)
    n = max(len(instruction_transition_constraints), len(padding))
    out = []
    for idx in range(n):
        z = e.constant(0)
        ic = instruction_transition_constraints[idx] if idx < len(instruction_transition_constraints) else z
        z = e.constant(0)
        pc = padding[idx] if idx < len(padding) else z
        out.append(ic * padding_row_deselector + pc * padding_row_selector)
    return out


# ---- table-linking constraints (processor.rs:2200-3052) ------------------------------------------
def log_derivative_accumulates_clk_next(e):
    return ((e.next_aux(A.ClockJumpDifferenceLookupServerLogDerivative)
             - e.cur_aux(A.ClockJumpDifferenceLookupServerLogDerivative))
            * (e.challenge(CH.ClockJumpDifferenceLookupIndeterminate) - e.next_main(C.CLK))
            - e.next_main(C.ClockJumpDifferenceLookupMultiplicity))


def log_derivative_for_instruction_lookup_updates_correctly(e):
    one = lambda: e.constant(1)
    compressed_row = (e.challenge(CH.ProgramAddressWeight) * e.next_main(C.IP)
                      + e.challenge(CH.ProgramInstructionWeight) * e.next_main(C.CI)
                      + e.challenge(CH.ProgramNextInstructionWeight) * e.next_main(C.NIA))
    updates = ((e.next_aux(A.InstructionLookupClientLogDerivative) - e.cur_aux(A.InstructionLookupClientLogDerivative))
               * (e.challenge(CH.InstructionLookupIndeterminate) - compressed_row) - one())
    remains = e.next_aux(A.InstructionLookupClientLogDerivative) - e.cur_aux(A.InstructionLookupClientLogDerivative)
    return (one() - e.next_main(C.IsPadding)) * updates + e.next_main(C.IsPadding) * remains


def running_product_for_jump_stack_table_updates_correctly(e):
    compressed_row = (e.challenge(CH.JumpStackClkWeight) * e.next_main(C.CLK)
                      + e.challenge(CH.JumpStackCiWeight) * e.next_main(C.CI)
                      + e.challenge(CH.JumpStackJspWeight) * e.next_main(C.JSP)
                      + e.challenge(CH.JumpStackJsoWeight) * e.next_main(C.JSO)
                      + e.challenge(CH.JumpStackJsdWeight) * e.next_main(C.JSD))
    return (e.next_aux(A.JumpStackTablePermArg)
            - e.cur_aux(A.JumpStackTablePermArg) * (e.challenge(CH.JumpStackIndeterminate) - compressed_row))


def running_evaluation_hash_input_updates_correctly(e):
    one = lambda: e.constant(1)
    hash_deselector = instruction_deselector_next_row(e, "hash")
    merkle_step_deselector = instruction_deselector_next_row(e, "merkle_step")
    merkle_step_mem_deselector = instruction_deselector_next_row(e, "merkle_step_mem")
    hash_and_merkle_step_selector = ((e.next_main(C.CI) - e.constant(OPCODE["hash"]))
                                     * (e.next_main(C.CI) - e.constant(OPCODE["merkle_step"]))
                                     * (e.next_main(C.CI) - e.constant(OPCODE["merkle_step_mem"])))
    weights = [stack_weight(e, i) for i in range(10)]
    state_for_hash = [e.next_main(st(i)) for i in range(10)]
    compressed_hash_row = msum(w * s for w, s in zip(weights, state_for_hash))

    is_left_sibling = lambda: e.next_main(C.HV5)
    is_right_sibling = lambda: one() - e.next_main(C.HV5)
    mse = lambda l, r: is_right_sibling() * e.next_main(l) + is_left_sibling() * e.next_main(r)
    state_for_merkle_step = ([mse(st(i), hvc(i)) for i in range(5)] + [mse(hvc(i), st(i)) for i in range(5)])
    compressed_merkle_step_row = msum(w * s for w, s in zip(weights, state_for_merkle_step))

    def running_evaluation_updates_with(compressed_row):
        return (e.next_aux(A.HashInputEvalArg)
                - e.challenge(CH.HashInputIndeterminate) * e.cur_aux(A.HashInputEvalArg)
                - compressed_row)

    running_evaluation_remains = e.next_aux(A.HashInputEvalArg) - e.cur_aux(A.HashInputEvalArg)
    return (hash_and_merkle_step_selector * running_evaluation_remains
            + hash_deselector * running_evaluation_updates_with(compressed_hash_row)
            + merkle_step_deselector * running_evaluation_updates_with(compressed_merkle_step_row)
            + merkle_step_mem_deselector * running_evaluation_updates_with(compressed_merkle_step_row))


def running_evaluation_hash_digest_updates_correctly(e):
    hash_deselector = instruction_deselector_current_row(e, "hash")
    merkle_step_deselector = instruction_deselector_current_row(e, "merkle_step")
    merkle_step_mem_deselector = instruction_deselector_current_row(e, "merkle_step_mem")
    selector = ((e.cur_main(C.CI) - e.constant(OPCODE["hash"]))
                * (e.cur_main(C.CI) - e.constant(OPCODE["merkle_step"]))
                * (e.cur_main(C.CI) - e.constant(OPCODE["merkle_step_mem"])))
    weights = [stack_weight(e, i) for i in range(5)]
    state = [e.next_main(st(i)) for i in range(5)]
    compressed_row = msum(w * s for w, s in zip(weights, state))
    updates = (e.next_aux(A.HashDigestEvalArg)
               - e.challenge(CH.HashDigestIndeterminate) * e.cur_aux(A.HashDigestEvalArg)
               - compressed_row)
    remains = e.next_aux(A.HashDigestEvalArg) - e.cur_aux(A.HashDigestEvalArg)
    return (selector * remains
            + (hash_deselector + merkle_step_deselector + merkle_step_mem_deselector) * updates)


def running_evaluation_sponge_updates_correctly(e):
    sponge_init_deselector = instruction_deselector_current_row(e, "sponge_init")
    sponge_absorb_deselector = instruction_deselector_current_row(e, "sponge_absorb")
    sponge_absorb_mem_deselector = instruction_deselector_current_row(e, "sponge_absorb_mem")
    sponge_squeeze_deselector = instruction_deselector_current_row(e, "sponge_squeeze")
    selector = ((e.cur_main(C.CI) - e.constant(OPCODE["sponge_init"]))
                * (e.cur_main(C.CI) - e.constant(OPCODE["sponge_absorb"]))
                * (e.cur_main(C.CI) - e.constant(OPCODE["sponge_absorb_mem"]))
                * (e.cur_main(C.CI) - e.constant(OPCODE["sponge_squeeze"])))

    def weighted_sum(state):
        weights = [stack_weight(e, i) for i in range(10)]      # eager array map
        return msum(w * s for w, s in zip(weights, state))

    compressed_row_current = weighted_sum([e.cur_main(st(i)) for i in range(10)])
    compressed_row_next = weighted_sum([e.next_main(st(i)) for i in range(10)])

    updates_for_sponge_init = (e.next_aux(A.SpongeEvalArg)
                               - e.challenge(CH.SpongeIndeterminate) * e.cur_aux(A.SpongeEvalArg)
                               - e.challenge(CH.HashCIWeight) * e.cur_main(C.CI))
    updates_for_absorb = updates_for_sponge_init - compressed_row_current
    updates_for_squeeze = updates_for_sponge_init - compressed_row_next
    remains = e.next_aux(A.SpongeEvalArg) - e.cur_aux(A.SpongeEvalArg)

    stack_elements = [e.next_main(st(i)) for i in range(1, 5)]
    hv_elements = [e.cur_main(hvc(i)) for i in range(6)]
    compressed_row_absorb_mem = weighted_sum(stack_elements + hv_elements)
    updates_for_absorb_mem = (e.next_aux(A.SpongeEvalArg)
                              - e.challenge(CH.SpongeIndeterminate) * e.cur_aux(A.SpongeEvalArg)
                              - e.challenge(CH.HashCIWeight) * e.constant(OPCODE["sponge_absorb"])
                              - compressed_row_absorb_mem)
    return (selector * remains
            + sponge_init_deselector * updates_for_sponge_init
            + sponge_absorb_deselector * updates_for_absorb
            + sponge_absorb_mem_deselector * updates_for_absorb_mem
            + sponge_squeeze_deselector * updates_for_squeeze)


def log_derivative_with_u32_table_updates_correctly(e):
    one = lambda: e.constant(1)
    two_inverse = e.constant(pow(2, P - 2, P))
    desel = lambda n: instruction_deselector_current_row(e, n)
    split_deselector = desel("split")
    lt_deselector = desel("lt")
    and_deselector = desel("and")
    xor_deselector = desel("xor")
    pow_deselector = desel("pow")
    log_2_floor_deselector = desel("log_2_floor")
    div_mod_deselector = desel("div_mod")
    pop_count_deselector = desel("pop_count")
    merkle_step_deselector = desel("merkle_step")
    merkle_step_mem_deselector = desel("merkle_step_mem")

    running_sum = e.cur_aux(A.U32LookupClientLogDerivative)
    running_sum_next = e.next_aux(A.U32LookupClientLogDerivative)

    ch = e.challenge
    split_factor = (ch(CH.U32Indeterminate) - ch(CH.U32LhsWeight) * e.next_main(C.ST0)
                    - ch(CH.U32RhsWeight) * e.next_main(C.ST1) - ch(CH.U32CiWeight) * e.cur_main(C.CI))
    binop_factor = (ch(CH.U32Indeterminate) - ch(CH.U32LhsWeight) * e.cur_main(C.ST0)
                    - ch(CH.U32RhsWeight) * e.cur_main(C.ST1) - ch(CH.U32CiWeight) * e.cur_main(C.CI)
                    - ch(CH.U32ResultWeight) * e.next_main(C.ST0))
    xor_factor = (ch(CH.U32Indeterminate) - ch(CH.U32LhsWeight) * e.cur_main(C.ST0)
                  - ch(CH.U32RhsWeight) * e.cur_main(C.ST1)
                  - ch(CH.U32CiWeight) * e.constant(OPCODE["and"])
                  - ch(CH.U32ResultWeight) * (e.cur_main(C.ST0) + e.cur_main(C.ST1) - e.next_main(C.ST0))
                  * two_inverse)
    unop_factor = (ch(CH.U32Indeterminate) - ch(CH.U32LhsWeight) * e.cur_main(C.ST0)
                   - ch(CH.U32CiWeight) * e.cur_main(C.CI) - ch(CH.U32ResultWeight) * e.next_main(C.ST0))
    div_mod_factor_for_lt = (ch(CH.U32Indeterminate) - ch(CH.U32LhsWeight) * e.next_main(C.ST0)
                             - ch(CH.U32RhsWeight) * e.cur_main(C.ST1)
                             - ch(CH.U32CiWeight) * e.constant(OPCODE["lt"]) - ch(CH.U32ResultWeight))
    div_mod_factor_for_range_check = (ch(CH.U32Indeterminate) - ch(CH.U32LhsWeight) * e.cur_main(C.ST0)
                                      - ch(CH.U32RhsWeight) * e.next_main(C.ST1)
                                      - ch(CH.U32CiWeight) * e.constant(OPCODE["split"]))
    merkle_step_range_check_factor = (ch(CH.U32Indeterminate) - ch(CH.U32LhsWeight) * e.cur_main(C.ST5)
                                      - ch(CH.U32RhsWeight) * e.next_main(C.ST5)
                                      - ch(CH.U32CiWeight) * e.constant(OPCODE["split"]))

    absorbs_split = (running_sum_next - running_sum) * split_factor - one()
    absorbs_binop = (running_sum_next - running_sum) * binop_factor - one()
    absorbs_xor = (running_sum_next - running_sum) * xor_factor - one()
    absorbs_unop = (running_sum_next - running_sum) * unop_factor - one()
    absorbs_merkle_step = (running_sum_next - running_sum) * merkle_step_range_check_factor - one()

    split_summand = split_deselector * absorbs_split
    lt_summand = lt_deselector * absorbs_binop
    and_summand = and_deselector * absorbs_binop
    xor_summand = xor_deselector * absorbs_xor
    pow_summand = pow_deselector * absorbs_binop
    log_2_floor_summand = log_2_floor_deselector * absorbs_unop
    div_mod_summand = div_mod_deselector * ((running_sum_next - running_sum) * div_mod_factor_for_lt
                                            * div_mod_factor_for_range_check
                                            - div_mod_factor_for_lt - div_mod_factor_for_range_check)
    pop_count_summand = pop_count_deselector * absorbs_unop
    merkle_step_summand = merkle_step_deselector * absorbs_merkle_step
    merkle_step_mem_summand = merkle_step_mem_deselector * absorbs_merkle_step
    no_update_summand = (one() - e.cur_main(C.IB2)) * (running_sum_next - running_sum)

    return (split_summand + lt_summand + and_summand + xor_summand + pow_summand + log_2_floor_summand
            + div_mod_summand + pop_count_summand + merkle_step_summand + merkle_step_mem_summand
            + no_update_summand)


def transition(b):
    e = Env(b)
    clk_increases_by_1 = e.next_main(C.CLK) - e.cur_main(C.CLK) - e.constant(1)
    is_padding_is_0_or_does_not_change = e.cur_main(C.IsPadding) * (e.next_main(C.IsPadding) - e.cur_main(C.IsPadding))

    all_instruction_deselectors = [instruction_deselector_current_row(e, instr) for instr in ALL_INSTRUCTIONS]
    acc = e.constant(0)
    for d in all_instruction_deselectors:
        acc = acc + d
    exactly_one = acc - e.constant(1)
    instruction_independent_constraints = [clk_increases_by_1, is_padding_is_0_or_does_not_change, exactly_one]

    all_tc = [transition_constraints_for_instruction(e, instr) for instr in ALL_INSTRUCTIONS]
    deselected = combine_instruction_constraints_with_deselectors(e, all_instruction_deselectors, all_tc)
    doubly_deselected = combine_transition_constraints_with_padding_constraints(e, deselected)

    table_linking_constraints = [
        log_derivative_accumulates_clk_next(e),
        log_derivative_for_instruction_lookup_updates_correctly(e),
        running_product_for_jump_stack_table_updates_correctly(e),
        running_evaluation_hash_input_updates_correctly(e),
        running_evaluation_hash_digest_updates_correctly(e),
        running_evaluation_sponge_updates_correctly(e),
        log_derivative_with_u32_table_updates_correctly(e),
    ]
    return instruction_independent_constraints + doubly_deselected + table_linking_constraints
