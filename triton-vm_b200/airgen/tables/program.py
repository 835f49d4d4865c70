"""Program table AIR — restates triton-air/src/table/program.rs:28-276 statement by statement
(leaf-creation order is significant for node ids)."""
from ..columns import MAIN, AUX, CH, Env

C, A = MAIN["program"], AUX["program"]
RATE = 10  # Tip5::RATE


def initial(b):
    e = Env(b)
    address = e.main(C.Address)
    instruction = e.main(C.Instruction)
    index_in_chunk = e.main(C.IndexInChunk)
    is_hash_input_padding = e.main(C.IsHashInputPadding)
    instruction_lookup_log_derivative = e.aux(A.InstructionLookupServerLogDerivative)
    prepare_chunk_running_evaluation = e.aux(A.PrepareChunkRunningEvaluation)
    send_chunk_running_evaluation = e.aux(A.SendChunkRunningEvaluation)
    lookup_arg_initial = e.x_constant(0)
    eval_arg_initial = e.x_constant(1)
    prepare_chunk_indeterminate = e.challenge(CH.ProgramAttestationPrepareChunkIndeterminate)
    first_address_is_zero = address
    index_in_chunk_is_zero = index_in_chunk
    hash_input_padding_indicator_is_zero = is_hash_input_padding
    log_derivative_init = instruction_lookup_log_derivative - lookup_arg_initial
    prepare_chunk_absorbed_first = (prepare_chunk_running_evaluation
                                    - eval_arg_initial * prepare_chunk_indeterminate - instruction)
    send_chunk_default_initial = send_chunk_running_evaluation - eval_arg_initial
    return [first_address_is_zero, index_in_chunk_is_zero, hash_input_padding_indicator_is_zero,
            log_derivative_init, prepare_chunk_absorbed_first, send_chunk_default_initial]


def consistency(b):
    e = Env(b)
    one = e.constant(1)
    max_index_in_chunk = e.constant(RATE - 1)
    index_in_chunk = e.main(C.IndexInChunk)
    max_minus_index_in_chunk_inv = e.main(C.MaxMinusIndexInChunkInv)
    is_hash_input_padding = e.main(C.IsHashInputPadding)
    is_table_padding = e.main(C.IsTablePadding)
    max_minus_index_in_chunk = max_index_in_chunk - index_in_chunk
    c0 = (one - max_minus_index_in_chunk * max_minus_index_in_chunk_inv) * max_minus_index_in_chunk_inv
    c1 = (one - max_minus_index_in_chunk * max_minus_index_in_chunk_inv) * max_minus_index_in_chunk
    is_hash_input_padding_is_bit = is_hash_input_padding * (is_hash_input_padding - one)
    is_table_padding_is_bit = is_table_padding * (is_table_padding - one)
    table_padding_implies_hash_input_padding = is_table_padding * (one - is_hash_input_padding)
    return [c0, c1, is_hash_input_padding_is_bit, is_table_padding_is_bit, table_padding_implies_hash_input_padding]


def transition(b):
    e = Env(b)
    one = e.constant(1)
    rate_minus_one = e.constant(RATE - 1)
    prepare_chunk_indeterminate = e.challenge(CH.ProgramAttestationPrepareChunkIndeterminate)
    send_chunk_indeterminate = e.challenge(CH.ProgramAttestationSendChunkIndeterminate)

    address = e.cur_main(C.Address)
    instruction = e.cur_main(C.Instruction)
    lookup_multiplicity = e.cur_main(C.LookupMultiplicity)
    index_in_chunk = e.cur_main(C.IndexInChunk)
    max_minus_index_in_chunk_inv = e.cur_main(C.MaxMinusIndexInChunkInv)
    is_hash_input_padding = e.cur_main(C.IsHashInputPadding)
    is_table_padding = e.cur_main(C.IsTablePadding)
    log_derivative = e.cur_aux(A.InstructionLookupServerLogDerivative)
    prepare_chunk_running_evaluation = e.cur_aux(A.PrepareChunkRunningEvaluation)
    send_chunk_running_evaluation = e.cur_aux(A.SendChunkRunningEvaluation)

    address_next = e.next_main(C.Address)
    instruction_next = e.next_main(C.Instruction)
    index_in_chunk_next = e.next_main(C.IndexInChunk)
    max_minus_index_in_chunk_inv_next = e.next_main(C.MaxMinusIndexInChunkInv)
    is_hash_input_padding_next = e.next_main(C.IsHashInputPadding)
    is_table_padding_next = e.next_main(C.IsTablePadding)
    log_derivative_next = e.next_aux(A.InstructionLookupServerLogDerivative)
    prepare_chunk_running_evaluation_next = e.next_aux(A.PrepareChunkRunningEvaluation)
    send_chunk_running_evaluation_next = e.next_aux(A.SendChunkRunningEvaluation)

    address_increases_by_one = address_next - (address + one)
    is_table_padding_is_0_or_remains_unchanged = is_table_padding * (is_table_padding_next - is_table_padding)

    index_in_chunk_cycles_correctly = ((one - max_minus_index_in_chunk_inv * (rate_minus_one - index_in_chunk))
                                       * index_in_chunk_next
                                       + max_minus_index_in_chunk_inv * (index_in_chunk_next - index_in_chunk - one))

    hash_input_indicator_is_0_or_remains_unchanged = is_hash_input_padding * (is_hash_input_padding_next - one)

    first_hash_input_padding_is_1 = ((is_hash_input_padding - one) * is_hash_input_padding_next
                                     * (instruction_next - one))

    hash_input_padding_is_0_after_the_first_1 = is_hash_input_padding * instruction_next

    next_row_is_table_padding_row = is_table_padding_next - one
    table_padding_starts = (is_hash_input_padding
                            * (one - max_minus_index_in_chunk_inv * (rate_minus_one - index_in_chunk))
                            * next_row_is_table_padding_row)

    log_derivative_remains = log_derivative_next - log_derivative
    compressed_row = (e.challenge(CH.ProgramAddressWeight) * address
                      + e.challenge(CH.ProgramInstructionWeight) * instruction
                      + e.challenge(CH.ProgramNextInstructionWeight) * instruction_next)

    indeterminate = e.challenge(CH.InstructionLookupIndeterminate)
    log_derivative_updates = ((log_derivative_next - log_derivative) * (indeterminate - compressed_row)
                              - lookup_multiplicity)
    log_derivative_updates_iff_not_padding = ((one - is_hash_input_padding) * log_derivative_updates
                                              + is_hash_input_padding * log_derivative_remains)

    prepare_absorbs_next = (prepare_chunk_running_evaluation_next
                            - prepare_chunk_indeterminate * prepare_chunk_running_evaluation
                            - instruction_next)
    prepare_resets_and_absorbs_next = (prepare_chunk_running_evaluation_next
                                       - prepare_chunk_indeterminate - instruction_next)
    index_in_chunk_is_max = rate_minus_one - index_in_chunk
    index_in_chunk_is_not_max = one - max_minus_index_in_chunk_inv * (rate_minus_one - index_in_chunk)
    prepare_chunk_constraint = (index_in_chunk_is_max * prepare_absorbs_next
                                + index_in_chunk_is_not_max * prepare_resets_and_absorbs_next)

    send_absorbs_next_chunk = (send_chunk_running_evaluation_next
                               - send_chunk_indeterminate * send_chunk_running_evaluation
                               - prepare_chunk_running_evaluation_next)
    send_does_not_change = send_chunk_running_evaluation_next - send_chunk_running_evaluation
    index_in_chunk_next_is_max = rate_minus_one - index_in_chunk_next
    index_in_chunk_next_is_not_max = one - max_minus_index_in_chunk_inv_next * index_in_chunk_next_is_max

    send_chunk_constraint = (send_absorbs_next_chunk * next_row_is_table_padding_row * index_in_chunk_next_is_not_max
                             + send_does_not_change * is_table_padding_next
                             + send_does_not_change * index_in_chunk_next_is_max)

    return [address_increases_by_one, is_table_padding_is_0_or_remains_unchanged, index_in_chunk_cycles_correctly,
            hash_input_indicator_is_0_or_remains_unchanged, first_hash_input_padding_is_1,
            hash_input_padding_is_0_after_the_first_1, table_padding_starts, log_derivative_updates_iff_not_padding,
            prepare_chunk_constraint, send_chunk_constraint]


def terminal(b):
    e = Env(b)
    index_in_chunk = e.main(C.IndexInChunk)
    is_hash_input_padding = e.main(C.IsHashInputPadding)
    is_table_padding = e.main(C.IsTablePadding)
    hash_input_padding_is_one = is_hash_input_padding - e.constant(1)
    index_in_chunk_is_max_or_row_is_padding_row = ((index_in_chunk - e.constant(RATE - 1))
                                                   * (is_table_padding - e.constant(1)))
    return [hash_input_padding_is_one, index_in_chunk_is_max_or_row_is_padding_row]
