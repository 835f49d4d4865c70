"""Lookup table AIR — restates triton-air/src/table/lookup.rs:42-183."""
from ..columns import MAIN, AUX, CH, Env

C, A = MAIN["lookup"], AUX["lookup"]


def initial(b):
    e = Env(b)
    lookup_input = e.main(C.LookIn)
    lookup_output = e.main(C.LookOut)
    lookup_multiplicity = e.main(C.LookupMultiplicity)
    cascade_ld = e.aux(A.CascadeTableServerLogDerivative)
    public_ea = e.aux(A.PublicEvaluationArgument)
    lookup_input_is_0 = lookup_input
    lookup_argument_default_initial = e.x_constant(0)
    cascade_table_indeterminate = e.challenge(CH.CascadeLookupIndeterminate)
    compressed_row = lookup_output * e.challenge(CH.LookupTableOutputWeight)
    c1 = ((cascade_ld - lookup_argument_default_initial) * (cascade_table_indeterminate - compressed_row)
          - lookup_multiplicity)
    eval_argument_default_initial = e.x_constant(1)
    public_indeterminate = e.challenge(CH.LookupTablePublicIndeterminate)
    c2 = public_ea - eval_argument_default_initial * public_indeterminate - lookup_output
    return [lookup_input_is_0, c1, c2]


def consistency(b):
    e = Env(b)
    padding_is_0_or_1 = e.main(C.IsPadding) * (e.constant(1) - e.main(C.IsPadding))
    return [padding_is_0_or_1]


def transition(b):
    e = Env(b)
    one = lambda: e.constant(1)
    lookup_input = e.cur_main(C.LookIn)
    is_padding = e.cur_main(C.IsPadding)
    cascade_ld = e.cur_aux(A.CascadeTableServerLogDerivative)
    public_ea = e.cur_aux(A.PublicEvaluationArgument)

    lookup_input_next = e.next_main(C.LookIn)
    lookup_output_next = e.next_main(C.LookOut)
    lookup_multiplicity_next = e.next_main(C.LookupMultiplicity)
    is_padding_next = e.next_main(C.IsPadding)
    cascade_ld_next = e.next_aux(A.CascadeTableServerLogDerivative)
    public_ea_next = e.next_aux(A.PublicEvaluationArgument)

    c0 = is_padding * (one() - is_padding_next)
    t0 = is_padding_next * lookup_input_next
    t1 = (one() - is_padding_next) * (lookup_input_next - lookup_input - one())
    c1 = t0 + t1

    cascade_table_indeterminate = e.challenge(CH.CascadeLookupIndeterminate)
    compressed_row = (lookup_input_next * e.challenge(CH.LookupTableInputWeight)
                      + lookup_output_next * e.challenge(CH.LookupTableOutputWeight))
    ld_remains = cascade_ld_next - cascade_ld
    ld_updates = (cascade_ld_next - cascade_ld) * (cascade_table_indeterminate - compressed_row) - lookup_multiplicity_next
    c2 = (one() - is_padding_next) * ld_updates + is_padding_next * ld_remains

    public_indeterminate = e.challenge(CH.LookupTablePublicIndeterminate)
    pea_remains = public_ea_next - public_ea
    pea_updates = public_ea_next - public_ea * public_indeterminate - lookup_output_next
    c3 = (one() - is_padding_next) * pea_updates + is_padding_next * pea_remains
    return [c0, c1, c2, c3]


def terminal(b):
    e = Env(b)
    return [e.aux(A.PublicEvaluationArgument) - e.challenge(CH.LookupTablePublicTerminal)]
