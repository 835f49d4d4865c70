"""Grand cross-table argument — restates triton-air/src/cross_table_argument.rs:107-213."""
from .columns import AUX, CH, Env


def initial(b): return []
def consistency(b): return []
def transition(b): return []


def terminal(b):
    e = Env(b)
    prog, proc, ops, ram, js, hsh, casc, lk, u32 = (AUX[k] for k in
        ("program", "processor", "op_stack", "ram", "jump_stack", "hash", "cascade", "lookup", "u32"))
    a = e.aux
    program_attestation = a(prog.SendChunkRunningEvaluation) - a(hsh.ReceiveChunkRunningEvaluation)
    input_to_processor = e.challenge(CH.StandardInputTerminal) - a(proc.InputTableEvalArg)
    processor_to_output = a(proc.OutputTableEvalArg) - e.challenge(CH.StandardOutputTerminal)
    instruction_lookup = a(proc.InstructionLookupClientLogDerivative) - a(prog.InstructionLookupServerLogDerivative)
    processor_to_op_stack = a(proc.OpStackTablePermArg) - a(ops.RunningProductPermArg)
    processor_to_ram = a(proc.RamTablePermArg) - a(ram.RunningProductPermArg)
    processor_to_jump_stack = a(proc.JumpStackTablePermArg) - a(js.RunningProductPermArg)
    hash_input = a(proc.HashInputEvalArg) - a(hsh.HashInputRunningEvaluation)
    hash_digest = a(hsh.HashDigestRunningEvaluation) - a(proc.HashDigestEvalArg)
    sponge = a(proc.SpongeEvalArg) - a(hsh.SpongeRunningEvaluation)
    hash_to_cascade = a(casc.HashTableServerLogDerivative)
    for i in range(4):
        for limb in ("Highest", "MidHigh", "MidLow", "Lowest"):
            hash_to_cascade = hash_to_cascade - a(getattr(hsh, f"CascadeState{i}{limb}ClientLogDerivative"))
    cascade_to_lookup = a(casc.LookupTableClientLogDerivative) - a(lk.CascadeTableServerLogDerivative)
    processor_to_u32 = a(proc.U32LookupClientLogDerivative) - a(u32.LookupServerLogDerivative)
    clock_jump_difference_lookup = (a(proc.ClockJumpDifferenceLookupServerLogDerivative)
                                    - a(ops.ClockJumpDifferenceLookupClientLogDerivative)
                                    - a(ram.ClockJumpDifferenceLookupClientLogDerivative)
                                    - a(js.ClockJumpDifferenceLookupClientLogDerivative))
    return [program_attestation, input_to_processor, processor_to_output, instruction_lookup, processor_to_op_stack,
            processor_to_ram, processor_to_jump_stack, hash_input, hash_digest, sponge, hash_to_cascade,
            cascade_to_lookup, processor_to_u32, clock_jump_difference_lookup]
