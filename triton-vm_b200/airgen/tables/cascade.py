"""Cascade table AIR — restates triton-air/src/table/cascade.rs:33-214."""
from ..columns import MAIN, AUX, CH, Env

C, A = MAIN["cascade"], AUX["cascade"]


def initial(b):
    e = Env(b)
    one = lambda: e.constant(1)
    two = lambda: e.constant(2)
    two_pow_8 = e.constant(1 << 8)
    lookup_arg_default_initial = e.x_constant(0)

    is_padding = e.main(C.IsPadding)
    look_in_hi = e.main(C.LookInHi)
    look_in_lo = e.main(C.LookInLo)
    look_out_hi = e.main(C.LookOutHi)
    look_out_lo = e.main(C.LookOutLo)
    lookup_multiplicity = e.main(C.LookupMultiplicity)
    hash_ld = e.aux(A.HashTableServerLogDerivative)
    lookup_ld = e.aux(A.LookupTableClientLogDerivative)

    hash_indeterminate = e.challenge(CH.HashCascadeLookupIndeterminate)
    hash_input_weight = e.challenge(CH.HashCascadeLookInWeight)
    hash_output_weight = e.challenge(CH.HashCascadeLookOutWeight)
    lookup_indeterminate = e.challenge(CH.CascadeLookupIndeterminate)
    lookup_input_weight = e.challenge(CH.LookupTableInputWeight)
    lookup_output_weight = e.challenge(CH.LookupTableOutputWeight)

    compressed_row_hash = (hash_input_weight * (two_pow_8 * look_in_hi + look_in_lo)
                           + hash_output_weight * (two_pow_8 * look_out_hi + look_out_lo))
    hash_ld_is_default_initial = hash_ld - lookup_arg_default_initial
    hash_ld_accumulated_first_row = ((hash_ld - lookup_arg_default_initial)
                                     * (hash_indeterminate - compressed_row_hash) - lookup_multiplicity)
    c0 = (one() - is_padding) * hash_ld_accumulated_first_row + is_padding * hash_ld_is_default_initial

    compressed_row_lo = lookup_input_weight * look_in_lo + lookup_output_weight * look_out_lo
    compressed_row_hi = lookup_input_weight * look_in_hi + lookup_output_weight * look_out_hi
    lookup_ld_is_default_initial = lookup_ld - lookup_arg_default_initial
    lookup_ld_accumulated_first_row = ((lookup_ld - lookup_arg_default_initial)
                                       * (lookup_indeterminate - compressed_row_lo)
                                       * (lookup_indeterminate - compressed_row_hi)
                                       - two() * lookup_indeterminate + compressed_row_lo + compressed_row_hi)
    c1 = (one() - is_padding) * lookup_ld_accumulated_first_row + is_padding * lookup_ld_is_default_initial
    return [c0, c1]


def consistency(b):
    e = Env(b)
    one = e.constant(1)
    is_padding = e.main(C.IsPadding)
    return [is_padding * (one - is_padding)]


def transition(b):
    e = Env(b)
    one = e.constant(1)
    two = e.constant(2)
    two_pow_8 = e.constant(1 << 8)

    is_padding = e.cur_main(C.IsPadding)
    hash_ld = e.cur_aux(A.HashTableServerLogDerivative)
    lookup_ld = e.cur_aux(A.LookupTableClientLogDerivative)

    is_padding_next = e.next_main(C.IsPadding)
    look_in_hi_next = e.next_main(C.LookInHi)
    look_in_lo_next = e.next_main(C.LookInLo)
    look_out_hi_next = e.next_main(C.LookOutHi)
    look_out_lo_next = e.next_main(C.LookOutLo)
    lookup_multiplicity_next = e.next_main(C.LookupMultiplicity)
    hash_ld_next = e.next_aux(A.HashTableServerLogDerivative)
    lookup_ld_next = e.next_aux(A.LookupTableClientLogDerivative)

    hash_indeterminate = e.challenge(CH.HashCascadeLookupIndeterminate)
    hash_input_weight = e.challenge(CH.HashCascadeLookInWeight)
    hash_output_weight = e.challenge(CH.HashCascadeLookOutWeight)
    lookup_indeterminate = e.challenge(CH.CascadeLookupIndeterminate)
    lookup_input_weight = e.challenge(CH.LookupTableInputWeight)
    lookup_output_weight = e.challenge(CH.LookupTableOutputWeight)

    c0 = is_padding * (one - is_padding_next)

    compressed_next_row_hash = (hash_input_weight * (two_pow_8 * look_in_hi_next + look_in_lo_next)
                                + hash_output_weight * (two_pow_8 * look_out_hi_next + look_out_lo_next))
    hash_ld_remains = hash_ld_next - hash_ld
    hash_ld_accumulates = ((hash_ld_next - hash_ld) * (hash_indeterminate - compressed_next_row_hash)
                           - lookup_multiplicity_next)
    c1 = (one - is_padding_next) * hash_ld_accumulates + is_padding_next * hash_ld_remains

    compressed_row_lo_next = lookup_input_weight * look_in_lo_next + lookup_output_weight * look_out_lo_next
    compressed_row_hi_next = lookup_input_weight * look_in_hi_next + lookup_output_weight * look_out_hi_next
    lookup_ld_remains = lookup_ld_next - lookup_ld
    lookup_ld_accumulates = ((lookup_ld_next - lookup_ld)
                             * (lookup_indeterminate - compressed_row_lo_next)
                             * (lookup_indeterminate - compressed_row_hi_next)
                             - two * lookup_indeterminate + compressed_row_lo_next + compressed_row_hi_next)
    c2 = (one - is_padding_next) * lookup_ld_accumulates + is_padding_next * lookup_ld_remains
    return [c0, c1, c2]


def terminal(b):
    return []
