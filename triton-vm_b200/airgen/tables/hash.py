"""Hash table AIR — restates triton-air/src/table/hash.rs:47-1350 statement by statement.

Rust iterator adaptors (`map`) are lazy: where the reference builds a `map` and only consumes
it later (`constraints.extend(...)`, hash.rs:775-776), the nodes are created at the point of
consumption; this restatement keeps that order because node ids depend on it."""
from ..columns import MAIN, AUX, CH, Env
from ..isa import OPCODE

C, A = MAIN["hash"], AUX["hash"]
P = (1 << 64) - (1 << 32) + 1
NUM_ROUNDS = 5
RATE = 10
DIGEST_LEN = 5
STATE_SIZE = 16
POWER_MAP_EXPONENT = 7
MONTGOMERY_MODULUS = (1 << 64) % P                      # hash.rs:29-30
MODE = {"ProgramHashing": 1, "Sponge": 2, "Hash": 3, "Pad": 0}   # hash.rs:1373-1382
MODE_ITER = ["ProgramHashing", "Sponge", "Hash", "Pad"]          # enum order, hash.rs:1354-1371
MDS_FIRST_COLUMN = [61402, 1108, 28750, 33823, 7454, 43244, 53865, 12034,
                    56951, 27521, 41351, 40901, 12021, 59689, 26798, 17845]

LIMBS = ("Highest", "MidHigh", "MidLow", "Lowest")


def lk_in(i, limb): return getattr(C, f"State{i}{limb}LkIn")
def lk_out(i, limb): return getattr(C, f"State{i}{limb}LkOut")
def casc(i, limb): return getattr(A, f"CascadeState{i}{limb}ClientLogDerivative")
def state_col(i): return getattr(C, f"State{i}")
def const_col(i): return getattr(C, f"Constant{i}")


def tip5_round_constants():
    """tip5::ROUND_CONSTANTS (canonical values): tips/tip-0005/tip-0005.md:72."""
    import blake3
    rinv = pow(MONTGOMERY_MODULUS, P - 2, P)
    return [int.from_bytes(blake3.blake3(b"Tip5" + bytes([i])).digest()[:16], "little") % P * rinv % P
            for i in range(STATE_SIZE * NUM_ROUNDS)]


def mds_matrix_entry(row, col):
    return MDS_FIRST_COLUMN[(STATE_SIZE + row - col) % STATE_SIZE]


def re_compose_16_bit_limbs(e, highest, mid_high, mid_low, lowest):
    montgomery_modulus_inv = e.constant(pow(MONTGOMERY_MODULUS, P - 2, P))
    s = highest * e.constant(1 << 48) + mid_high * e.constant(1 << 32) + mid_low * e.constant(1 << 16) + lowest
    return s * montgomery_modulus_inv


def round_number_deselector(e, rn, to_deselect):
    first_factor = e.constant(1) if to_deselect == 0 else rn
    acc = first_factor
    for r in range(1, NUM_ROUNDS + 1):
        if r == to_deselect:
            continue
        acc = acc * (rn - e.constant(r))
    return acc


def select_mode(e, mode_node, mode):
    return mode_node - e.constant(MODE[mode])


def mode_deselector(e, mode_node, mode_to_deselect):
    acc = e.constant(1)
    for m in MODE_ITER:
        if m == mode_to_deselect:
            continue
        acc = acc * (mode_node - e.constant(MODE[m]))
    return acc


def instruction_deselector(e, ci_node, to_deselect):
    acc = e.constant(1)
    for instr in ["hash", "sponge_init", "sponge_absorb", "sponge_squeeze"]:
        if instr == to_deselect:
            continue
        acc = acc * (ci_node - e.constant(OPCODE[instr]))
    return acc


def re_compose_states_0_through_3_before_lookup(e, inp):
    out = []
    for i in range(4):
        out.append(re_compose_16_bit_limbs(e, inp(lk_in(i, "Highest")), inp(lk_in(i, "MidHigh")),
                                           inp(lk_in(i, "MidLow")), inp(lk_in(i, "Lowest"))))
    return out


def tip5_constraints_as_circuits(e):
    after_lookup = []
    for i in range(4):
        after_lookup.append(re_compose_16_bit_limbs(e, e.cur_main(lk_out(i, "Highest")), e.cur_main(lk_out(i, "MidHigh")),
                                                    e.cur_main(lk_out(i, "MidLow")), e.cur_main(lk_out(i, "Lowest"))))
    before_power_map = [e.cur_main(state_col(i)) for i in range(4, 16)]
    acc = list(before_power_map)
    for _ in range(1, POWER_MAP_EXPONENT):
        for i in range(len(acc)):
            acc[i] = acc[i] * before_power_map[i]
    after_sbox = after_lookup + acc

    zero = e.constant(0)
    after_mds = [zero] * STATE_SIZE
    for row in range(STATE_SIZE):
        for col in range(STATE_SIZE):
            matrix_entry = e.constant(mds_matrix_entry(row, col))
            after_mds[row] = after_mds[row] + matrix_entry * after_sbox[col]

    round_constants = [e.cur_main(const_col(i)) for i in range(16)]
    after_rc = [st + rc for st, rc in zip(after_mds, round_constants)]

    s0123 = re_compose_states_0_through_3_before_lookup(e, e.next_main)
    state_next = s0123 + [e.next_main(state_col(i)) for i in range(4, 16)]
    round_number_next = e.next_main(C.RoundNumber)
    update = [round_number_next * (se - sn) for se, sn in zip(after_rc, state_next)]
    return state_next, update


def cascade_log_derivative_update_circuit(e, look_in_column, look_out_column, cascade_column):
    cascade_indeterminate = e.challenge(CH.HashCascadeLookupIndeterminate)
    look_in_weight = e.challenge(CH.HashCascadeLookInWeight)
    look_out_weight = e.challenge(CH.HashCascadeLookOutWeight)

    ci_next = e.next_main(C.CI)
    mode_next = e.next_main(C.Mode)
    round_number_next = e.next_main(C.RoundNumber)
    cld = e.cur_aux(cascade_column)
    cld_next = e.next_aux(cascade_column)

    compressed_row = look_in_weight * e.next_main(look_in_column) + look_out_weight * e.next_main(look_out_column)
    cld_remains = cld_next - cld
    cld_updates = (cld_next - cld) * (cascade_indeterminate - compressed_row) - e.constant(1)

    t0 = (select_mode(e, mode_next, "Pad") * (round_number_next - e.constant(NUM_ROUNDS))
          * (ci_next - e.constant(OPCODE["sponge_init"])))
    rn_next_is_not_num_rounds = round_number_deselector(e, round_number_next, NUM_ROUNDS)
    ci_next_is_not_sponge_init = instruction_deselector(e, ci_next, "sponge_init")
    next_row_is_padding_row = mode_deselector(e, mode_next, "Pad")
    return (t0 * cld_updates + rn_next_is_not_num_rounds * cld_remains
            + ci_next_is_not_sponge_init * cld_remains + next_row_is_padding_row * cld_remains)


def initial(b):
    e = Env(b)
    running_evaluation_initial = e.x_constant(1)
    lookup_arg_default_initial = e.x_constant(0)

    mode = e.main(C.Mode)
    re_hash_input = e.aux(A.HashInputRunningEvaluation)
    re_hash_digest = e.aux(A.HashDigestRunningEvaluation)
    re_sponge = e.aux(A.SpongeRunningEvaluation)
    re_receive_chunk = e.aux(A.ReceiveChunkRunningEvaluation)

    cascade_indeterminate = e.challenge(CH.HashCascadeLookupIndeterminate)
    look_in_weight = e.challenge(CH.HashCascadeLookInWeight)
    look_out_weight = e.challenge(CH.HashCascadeLookOutWeight)
    prepare_chunk_indeterminate = e.challenge(CH.ProgramAttestationPrepareChunkIndeterminate)
    receive_chunk_indeterminate = e.challenge(CH.ProgramAttestationSendChunkIndeterminate)

    s0123 = re_compose_states_0_through_3_before_lookup(e, e.main)
    state_rate_part = s0123 + [e.main(state_col(i)) for i in range(4, 10)]
    compressed_chunk = running_evaluation_initial
    for se in state_rate_part:
        compressed_chunk = compressed_chunk * prepare_chunk_indeterminate + se
    receive_chunk_init = (re_receive_chunk - receive_chunk_indeterminate * running_evaluation_initial
                          - compressed_chunk)

    def cascade_init(look_in_column, look_out_column, cascade_column):
        look_in = e.main(look_in_column)
        look_out = e.main(look_out_column)
        compressed_row = look_in_weight * look_in + look_out_weight * look_out
        cld = e.aux(cascade_column)
        return (cld - lookup_arg_default_initial) * (cascade_indeterminate - compressed_row) - e.constant(1)

    mode_is_program_hashing = select_mode(e, mode, "ProgramHashing")
    round_number_is_0 = e.main(C.RoundNumber)
    c_hi = re_hash_input - running_evaluation_initial
    c_hd = re_hash_digest - running_evaluation_initial
    c_sp = re_sponge - running_evaluation_initial

    out = [mode_is_program_hashing, round_number_is_0, c_hi, c_hd, c_sp, receive_chunk_init]
    for i in range(4):
        for limb in LIMBS:
            out.append(cascade_init(lk_in(i, limb), lk_out(i, limb), casc(i, limb)))
    return out


def consistency(b):
    e = Env(b)
    opcode = lambda n: e.constant(OPCODE[n])
    mode = e.main(C.Mode)
    ci = e.main(C.CI)
    round_number = e.main(C.RoundNumber)

    ci_is_hash = ci - opcode("hash")
    ci_is_sponge_init = ci - opcode("sponge_init")
    ci_is_sponge_absorb = ci - opcode("sponge_absorb")
    ci_is_sponge_squeeze = ci - opcode("sponge_squeeze")

    mode_is_not_hash = mode_deselector(e, mode, "Hash")
    round_number_is_not_0 = round_number_deselector(e, round_number, 0)

    mode_is_a_valid_mode = mode_deselector(e, mode, "Pad") * select_mode(e, mode, "Pad")
    if_mode_is_not_sponge_then_ci_is_hash = select_mode(e, mode, "Sponge") * ci_is_hash
    if_mode_is_sponge_then_ci_is_a_sponge_instruction = (mode_deselector(e, mode, "Sponge") * ci_is_sponge_init
                                                         * ci_is_sponge_absorb * ci_is_sponge_squeeze)
    if_padding_mode_then_round_number_is_0 = mode_deselector(e, mode, "Pad") * round_number
    if_ci_is_sponge_init_then_ = ci_is_hash * ci_is_sponge_absorb * ci_is_sponge_squeeze
    if_ci_is_sponge_init_then_round_number_is_0 = if_ci_is_sponge_init_then_ * round_number

    # lazy: consumed by `constraints.extend` below (hash.rs:648-651, 775)
    def if_ci_is_sponge_init_then_rate_is_0():
        out = []
        for state_index in range(10, 16):
            state_element = e.main(state_col(state_index))
            out.append(if_ci_is_sponge_init_then_ * state_element)
        return out

    if_mode_is_hash_and_round_no_is_0_then_ = round_number_is_not_0 * mode_is_not_hash

    def states_10_through_15_are_1():   # lazy as well (hash.rs:653-658, 776)
        out = []
        for state_index in range(10, 16):
            state_element = e.main(state_col(state_index))
            out.append(if_mode_is_hash_and_round_no_is_0_then_ * (state_element - e.constant(1)))
        return out

    one = e.constant(1)
    two_pow_16 = e.constant(1 << 16)
    two_pow_32 = e.constant(1 << 32)
    hi_minus = []
    for i in range(4):
        hi_minus.append(two_pow_32 - one - e.main(lk_in(i, "Highest")) * two_pow_16 - e.main(lk_in(i, "MidHigh")))
    hi_inv = [e.main(getattr(C, f"State{i}Inv")) for i in range(4)]
    not_all_1s = [hi_minus[i] * hi_inv[i] - one for i in range(4)]
    inv_is_inv_or_is_zero = [not_all_1s[i] * hi_inv[i] for i in range(4)]
    inv_is_inv_or_hi_is_zero = [not_all_1s[i] * hi_minus[i] for i in range(4)]
    lo_limbs = [e.main(lk_in(i, "MidLow")) * two_pow_16 + e.main(lk_in(i, "Lowest")) for i in range(4)]
    hi_all_1_then_lo_all_0 = [not_all_1s[i] * lo_limbs[i] for i in range(4)]

    constraints = ([mode_is_a_valid_mode, if_mode_is_not_sponge_then_ci_is_hash,
                    if_mode_is_sponge_then_ci_is_a_sponge_instruction, if_padding_mode_then_round_number_is_0,
                    if_ci_is_sponge_init_then_round_number_is_0]
                   + inv_is_inv_or_is_zero + inv_is_inv_or_hi_is_zero + hi_all_1_then_lo_all_0)
    constraints += if_ci_is_sponge_init_then_rate_is_0()
    constraints += states_10_through_15_are_1()

    rcs = tip5_round_constants()
    for col_idx in range(STATE_SIZE):
        rc_column_circuit = e.main(const_col(col_idx))
        acc = e.constant(0)
        for round_idx in range(NUM_ROUNDS):
            round_constant = e.constant(rcs[STATE_SIZE * round_idx + col_idx])
            desel = round_number_deselector(e, round_number, round_idx)
            acc = acc + desel * (rc_column_circuit - round_constant)
        constraints.append(acc)
    return constraints


def transition(b):
    e = Env(b)
    opcode = lambda n: e.constant(OPCODE[n])
    opcode_hash = opcode("hash")
    opcode_sponge_init = opcode("sponge_init")
    opcode_sponge_absorb = opcode("sponge_absorb")
    opcode_sponge_squeeze = opcode("sponge_squeeze")

    running_evaluation_initial = e.x_constant(1)
    prepare_chunk_indeterminate = e.challenge(CH.ProgramAttestationPrepareChunkIndeterminate)
    receive_chunk_indeterminate = e.challenge(CH.ProgramAttestationSendChunkIndeterminate)
    compress_program_digest_indeterminate = e.challenge(CH.CompressProgramDigestIndeterminate)
    expected_program_digest = e.challenge(CH.CompressedProgramDigest)
    hash_input_eval_indeterminate = e.challenge(CH.HashInputIndeterminate)
    hash_digest_eval_indeterminate = e.challenge(CH.HashDigestIndeterminate)
    sponge_indeterminate = e.challenge(CH.SpongeIndeterminate)

    mode = e.cur_main(C.Mode)
    ci = e.cur_main(C.CI)
    round_number = e.cur_main(C.RoundNumber)
    re_receive_chunk = e.cur_aux(A.ReceiveChunkRunningEvaluation)
    re_hash_input = e.cur_aux(A.HashInputRunningEvaluation)
    re_hash_digest = e.cur_aux(A.HashDigestRunningEvaluation)
    re_sponge = e.cur_aux(A.SpongeRunningEvaluation)

    mode_next = e.next_main(C.Mode)
    ci_next = e.next_main(C.CI)
    round_number_next = e.next_main(C.RoundNumber)
    re_receive_chunk_next = e.next_aux(A.ReceiveChunkRunningEvaluation)
    re_hash_input_next = e.next_aux(A.HashInputRunningEvaluation)
    re_hash_digest_next = e.next_aux(A.HashDigestRunningEvaluation)
    re_sponge_next = e.next_aux(A.SpongeRunningEvaluation)

    s0123 = re_compose_states_0_through_3_before_lookup(e, e.cur_main)
    state_current = s0123 + [e.cur_main(state_col(i)) for i in range(4, 16)]

    state_next, hash_function_round_correctly_performs_update = tip5_constraints_as_circuits(e)

    state_weights = [e.challenge(getattr(CH, f"StackWeight{i}")) for i in range(16)]

    round_number_is_not_num_rounds = round_number_deselector(e, round_number, NUM_ROUNDS)
    c0 = round_number_is_not_num_rounds * round_number_next

    c1 = (select_mode(e, mode_next, "Pad") * (ci - opcode_sponge_init)
          * (round_number - e.constant(NUM_ROUNDS)) * (round_number_next - round_number - e.constant(1)))

    c2 = instruction_deselector(e, ci, "sponge_init") * round_number_next

    compressed_digest = running_evaluation_initial
    for de in state_current[:DIGEST_LEN]:
        compressed_digest = compressed_digest * compress_program_digest_indeterminate + de
    c4 = (mode_deselector(e, mode, "ProgramHashing") * select_mode(e, mode_next, "ProgramHashing")
          * (compressed_digest - expected_program_digest))

    c5 = (mode_deselector(e, mode, "ProgramHashing") * mode_deselector(e, mode_next, "Sponge")
          * (ci_next - opcode_sponge_init))

    c6 = (round_number - e.constant(NUM_ROUNDS)) * (ci - opcode_sponge_init) * (ci_next - ci)
    c7 = (round_number - e.constant(NUM_ROUNDS)) * (ci - opcode_sponge_init) * (mode_next - mode)

    c8 = (mode_deselector(e, mode, "Sponge") * select_mode(e, mode_next, "Sponge")
          * select_mode(e, mode_next, "Hash") * select_mode(e, mode_next, "Pad"))
    c9 = mode_deselector(e, mode, "Hash") * select_mode(e, mode_next, "Hash") * select_mode(e, mode_next, "Pad")
    c10 = mode_deselector(e, mode, "Pad") * select_mode(e, mode_next, "Pad")

    diff_capacity = [nx - cu for cu, nx in zip(state_current[RATE:], state_next[RATE:])]
    from ..circuit import msum
    randomized_sum_of_capacity_differences = msum(w * d for w, d in zip(state_weights[RATE:], diff_capacity))

    c11 = (round_number_deselector(e, round_number_next, 0) * select_mode(e, mode_next, "Hash")
           * select_mode(e, mode_next, "Pad") * (ci_next - opcode_sponge_init)
           * randomized_sum_of_capacity_differences)

    diff_state = [nx - cu for cu, nx in zip(state_current, state_next)]
    randomized_sum_of_state_differences = msum(w * d for w, d in zip(state_weights, diff_state))
    c12 = (round_number_deselector(e, round_number_next, 0)
           * instruction_deselector(e, ci_next, "sponge_squeeze") * randomized_sum_of_state_differences)

    re_hash_input_remains = re_hash_input_next - re_hash_input
    compressed_row_from_processor = msum(w * s for s, w in zip(state_next[:RATE], state_weights[:RATE]))
    re_hash_input_updates = (re_hash_input_next - hash_input_eval_indeterminate * re_hash_input
                             - compressed_row_from_processor)
    c13 = (round_number_deselector(e, round_number_next, 0) * mode_deselector(e, mode_next, "Hash")
           * re_hash_input_updates
           + round_number_next * re_hash_input_remains
           + (e.constant(MODE["Hash"]) - mode_next) * re_hash_input_remains)

    round_number_next_is_num_rounds = round_number_next - e.constant(NUM_ROUNDS)
    re_hash_digest_remains = re_hash_digest_next - re_hash_digest
    compressed_row_hash_digest = msum(w * s for s, w in zip(state_next[:DIGEST_LEN], state_weights[:DIGEST_LEN]))
    re_hash_digest_updates = (re_hash_digest_next - hash_digest_eval_indeterminate * re_hash_digest
                              - compressed_row_hash_digest)
    c14 = (round_number_deselector(e, round_number_next, NUM_ROUNDS) * mode_deselector(e, mode_next, "Hash")
           * re_hash_digest_updates
           + round_number_next_is_num_rounds * re_hash_digest_remains
           + select_mode(e, mode_next, "Hash") * re_hash_digest_remains)

    compressed_row_next = msum(w * s for w, s in zip(state_weights[:RATE], state_next[:RATE]))
    re_sponge_has_accumulated_ci = (re_sponge_next - sponge_indeterminate * re_sponge
                                    - e.challenge(CH.HashCIWeight) * ci_next)
    re_sponge_has_accumulated_next_row = re_sponge_has_accumulated_ci - compressed_row_next
    t0 = (round_number_deselector(e, round_number_next, 0) * (ci_next - opcode_hash)
          * re_sponge_has_accumulated_next_row)
    re_sponge_remains = re_sponge_next - re_sponge
    t1 = round_number_next * re_sponge_remains
    t2 = ((ci_next - opcode_sponge_init) * (ci_next - opcode_sponge_absorb) * (ci_next - opcode_sponge_squeeze)
          * re_sponge_remains)
    c15 = t0 + t1 + t2

    compressed_chunk = running_evaluation_initial
    for re_ in state_next[:RATE]:
        compressed_chunk = compressed_chunk * prepare_chunk_indeterminate + re_
    receive_absorbs = re_receive_chunk_next - receive_chunk_indeterminate * re_receive_chunk - compressed_chunk
    receive_remains = re_receive_chunk_next - re_receive_chunk
    c3 = (round_number_deselector(e, round_number_next, 0) * mode_deselector(e, mode_next, "ProgramHashing")
          * receive_absorbs
          + round_number_next * receive_remains
          + select_mode(e, mode_next, "ProgramHashing") * receive_remains)

    constraints = [c0, c1, c2, c3, c4, c5, c6, c7, c8, c9, c10, c11, c12, c13, c14, c15]
    for i in range(4):
        for limb in LIMBS:
            constraints.append(cascade_log_derivative_update_circuit(e, lk_in(i, limb), lk_out(i, limb), casc(i, limb)))
    return constraints + hash_function_round_correctly_performs_update


def terminal(b):
    e = Env(b)
    mode = e.main(C.Mode)
    round_number = e.main(C.RoundNumber)
    compress_program_digest_indeterminate = e.challenge(CH.CompressProgramDigestIndeterminate)
    expected_program_digest = e.challenge(CH.CompressedProgramDigest)
    max_round_number = e.constant(NUM_ROUNDS)
    s0123 = re_compose_states_0_through_3_before_lookup(e, e.main)
    state_4 = e.main(C.State4)
    program_digest = s0123 + [state_4]
    compressed_digest = e.x_constant(1)
    for de in program_digest:
        compressed_digest = compressed_digest * compress_program_digest_indeterminate + de
    c0 = mode_deselector(e, mode, "ProgramHashing") * (compressed_digest - expected_program_digest)
    c1 = (select_mode(e, mode, "Pad") * (e.main(C.CI) - e.constant(OPCODE["sponge_init"]))
          * (round_number - max_round_number))
    return [c0, c1]
