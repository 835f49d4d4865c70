"""Master-table column indices and challenge ids.
Restates triton-air/src/table_column.rs (enum order = column order), table.rs:29-98 (table
offsets inside the master tables) and challenge_id.rs (enum order = challenge index)."""


class _Enum:
    def __init__(self, names, start):
        self.names = list(names)
        self.start = start
        self.COUNT = len(self.names)
        for i, n in enumerate(self.names):
            setattr(self, n, start + i)     # master index

    def local(self, master_index):
        return master_index - self.start


def _mk(main_specs, aux_specs):
    out, start = {}, 0
    for key, names in main_specs:
        out[key] = _Enum(names.split(), start)
        start += out[key].COUNT
    n_main = start
    outa, start = {}, 0
    for key, names in aux_specs:
        outa[key] = _Enum(names.split(), start)
        start += outa[key].COUNT
    return out, n_main, outa, start


_MAIN = [
    ("program", "Address Instruction LookupMultiplicity IndexInChunk MaxMinusIndexInChunkInv IsHashInputPadding IsTablePadding"),
    ("processor", "CLK IsPadding IP CI NIA IB0 IB1 IB2 IB3 IB4 IB5 IB6 JSP JSO JSD ST0 ST1 ST2 ST3 ST4 ST5 ST6 ST7 ST8 ST9 ST10 "
                  "ST11 ST12 ST13 ST14 ST15 OpStackPointer HV0 HV1 HV2 HV3 HV4 HV5 ClockJumpDifferenceLookupMultiplicity"),
    ("op_stack", "CLK IB1ShrinkStack StackPointer FirstUnderflowElement"),
    ("ram", "CLK InstructionType RamPointer RamValue InverseOfRampDifference BezoutCoefficientPolynomialCoefficient0 "
            "BezoutCoefficientPolynomialCoefficient1"),
    ("jump_stack", "CLK CI JSP JSO JSD"),
    ("hash", "Mode CI RoundNumber "
             + " ".join(f"State{i}{p}LkIn" for i in range(4) for p in ("Highest", "MidHigh", "MidLow", "Lowest")) + " "
             + " ".join(f"State{i}{p}LkOut" for i in range(4) for p in ("Highest", "MidHigh", "MidLow", "Lowest")) + " "
             + " ".join(f"State{i}" for i in range(4, 16)) + " "
             + " ".join(f"State{i}Inv" for i in range(4)) + " "
             + " ".join(f"Constant{i}" for i in range(16))),
    ("cascade", "IsPadding LookInHi LookInLo LookOutHi LookOutLo LookupMultiplicity"),
    ("lookup", "IsPadding LookIn LookOut LookupMultiplicity"),
    ("u32", "CopyFlag Bits BitsMinus33Inv CI LHS LhsInv RHS RhsInv Result LookupMultiplicity"),
]
_AUX = [
    ("program", "InstructionLookupServerLogDerivative PrepareChunkRunningEvaluation SendChunkRunningEvaluation"),
    ("processor", "InputTableEvalArg OutputTableEvalArg InstructionLookupClientLogDerivative OpStackTablePermArg RamTablePermArg "
                  "JumpStackTablePermArg HashInputEvalArg HashDigestEvalArg SpongeEvalArg U32LookupClientLogDerivative "
                  "ClockJumpDifferenceLookupServerLogDerivative"),
    ("op_stack", "RunningProductPermArg ClockJumpDifferenceLookupClientLogDerivative"),
    ("ram", "RunningProductOfRAMP FormalDerivative BezoutCoefficient0 BezoutCoefficient1 RunningProductPermArg "
            "ClockJumpDifferenceLookupClientLogDerivative"),
    ("jump_stack", "RunningProductPermArg ClockJumpDifferenceLookupClientLogDerivative"),
    ("hash", "ReceiveChunkRunningEvaluation HashInputRunningEvaluation HashDigestRunningEvaluation SpongeRunningEvaluation "
             + " ".join(f"CascadeState{i}{p}ClientLogDerivative" for i in range(4) for p in ("Highest", "MidHigh", "MidLow", "Lowest"))),
    ("cascade", "HashTableServerLogDerivative LookupTableClientLogDerivative"),
    ("lookup", "CascadeTableServerLogDerivative PublicEvaluationArgument"),
    ("u32", "LookupServerLogDerivative"),
]

MAIN, NUM_MAIN_COLUMNS, AUX, NUM_AUX_COLUMNS = _mk(_MAIN, _AUX)
assert NUM_MAIN_COLUMNS == 149 and NUM_AUX_COLUMNS == 49   # arithmetization-overview.md:7-20
assert MAIN["hash"].COUNT == 67 and AUX["hash"].COUNT == 20

_CH = ("CompressProgramDigestIndeterminate StandardInputIndeterminate StandardOutputIndeterminate "
       "InstructionLookupIndeterminate HashInputIndeterminate HashDigestIndeterminate SpongeIndeterminate "
       "OpStackIndeterminate RamIndeterminate JumpStackIndeterminate U32Indeterminate "
       "ClockJumpDifferenceLookupIndeterminate RamTableBezoutRelationIndeterminate ProgramAddressWeight "
       "ProgramInstructionWeight ProgramNextInstructionWeight OpStackClkWeight OpStackIb1Weight OpStackPointerWeight "
       "OpStackFirstUnderflowElementWeight RamClkWeight RamPointerWeight RamValueWeight RamInstructionTypeWeight "
       "JumpStackClkWeight JumpStackCiWeight JumpStackJspWeight JumpStackJsoWeight JumpStackJsdWeight "
       "ProgramAttestationPrepareChunkIndeterminate ProgramAttestationSendChunkIndeterminate HashCIWeight "
       + " ".join(f"StackWeight{i}" for i in range(16)) + " "
       "HashCascadeLookupIndeterminate HashCascadeLookInWeight HashCascadeLookOutWeight CascadeLookupIndeterminate "
       "LookupTableInputWeight LookupTableOutputWeight LookupTablePublicIndeterminate U32LhsWeight U32RhsWeight "
       "U32CiWeight U32ResultWeight "
       "StandardInputTerminal StandardOutputTerminal LookupTablePublicTerminal CompressedProgramDigest").split()


class _Ch:
    pass


CH = _Ch()
for _i, _n in enumerate(_CH):
    setattr(CH, _n, _i)
NUM_CHALLENGES = len(_CH)
NUM_DERIVED_CHALLENGES = 4
assert NUM_CHALLENGES == 63 and NUM_CHALLENGES - NUM_DERIVED_CHALLENGES == 59   # stark.rs:374-376


class Env:
    """The closures every table definition opens with (e.g. program.rs:33-36,112-119)."""

    def __init__(self, b):
        self.b = b

    def challenge(self, c): return self.b.challenge(c)
    def constant(self, c): return self.b.b_constant(c)
    def x_constant(self, c): return self.b.x_constant(c)
    # single-row
    def main(self, col): return self.b.input(0, True, col)
    def aux(self, col): return self.b.input(0, False, col)
    # dual-row
    def cur_main(self, col): return self.b.input(0, True, col)
    def cur_aux(self, col): return self.b.input(0, False, col)
    def next_main(self, col): return self.b.input(1, True, col)
    def next_aux(self, col): return self.b.input(1, False, col)
