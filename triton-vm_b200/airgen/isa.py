"""The slice of triton-isa the AIR needs: instruction order, opcodes, instruction bits.
Restates triton-isa/src/instruction.rs:29-76 (ALL_INSTRUCTIONS order) and 315-364 (opcodes)."""

ALL_INSTRUCTIONS = [
    "pop", "push", "divine", "pick", "place", "dup", "swap", "halt", "nop", "skiz", "call", "return", "recurse",
    "recurse_or_return", "assert", "read_mem", "write_mem", "hash", "assert_vector", "sponge_init", "sponge_absorb",
    "sponge_absorb_mem", "sponge_squeeze", "add", "addi", "mul", "invert", "eq", "split", "lt", "and", "xor",
    "log_2_floor", "pow", "div_mod", "pop_count", "xx_add", "xx_mul", "x_invert", "xb_mul", "read_io", "write_io",
    "merkle_step", "merkle_step_mem", "b_horner_step", "x_horner_step",
]

OPCODE = {
    "pop": 3, "push": 1, "divine": 9, "pick": 17, "place": 25, "dup": 33, "swap": 41,
    "halt": 0, "nop": 8, "skiz": 2, "call": 49, "return": 16, "recurse": 24,
    "recurse_or_return": 32, "assert": 10, "read_mem": 57, "write_mem": 11, "hash": 18,
    "assert_vector": 26, "sponge_init": 40, "sponge_absorb": 34, "sponge_absorb_mem": 48,
    "sponge_squeeze": 56, "add": 42, "addi": 65, "mul": 50, "invert": 64, "eq": 58,
    "split": 4, "lt": 6, "and": 14, "xor": 22, "log_2_floor": 12, "pow": 30, "div_mod": 20,
    "pop_count": 28, "xx_add": 66, "xx_mul": 74, "x_invert": 72, "xb_mul": 82,
    "read_io": 73, "write_io": 19, "merkle_step": 36, "merkle_step_mem": 44,
    "b_horner_step": 80, "x_horner_step": 88,
}
assert len(ALL_INSTRUCTIONS) == 46 and set(ALL_INSTRUCTIONS) == set(OPCODE)
NUM_INSTRUCTION_BITS = 7


def ib(instr, bit):
    """Instruction::ib: bit `bit` of the opcode (instruction.rs `ib`)."""
    return (OPCODE[instr] >> bit) & 1
