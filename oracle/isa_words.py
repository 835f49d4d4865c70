"""Minimal Triton-assembly -> program-words encoder, enough for the Tip5 KAT
`program_hash_is_unchanged` (reference `triton-vm/src/stark.rs:4828-4838`).

Opcode numbers: `triton-isa/src/instruction.rs:315-364`; words = opcode followed
by the argument if there is one (`triton-isa/src/program.rs:373-385`,
`instruction.rs:590-607`); `call` takes the absolute word address of its label.
TEST INFRASTRUCTURE ONLY.
"""
from .field import P

OPCODES = {
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
HAS_ARG = {"pop", "push", "divine", "pick", "place", "dup", "swap", "call",
           "read_mem", "write_mem", "addi", "read_io", "write_io"}


def assemble(text):
    toks = []
    for line in text.splitlines():
        line = line.split("//")[0]
        toks += line.split()
    items, labels, addr = [], {}, 0
    i = 0
    while i < len(toks):
        t = toks[i]
        if t.endswith(":"):
            labels[t[:-1]] = addr
            i += 1
            continue
        if t in HAS_ARG:
            items.append((t, toks[i + 1]))
            addr += 2
            i += 2
        else:
            items.append((t, None))
            addr += 1
            i += 1
    words = []
    for name, arg in items:
        words.append(OPCODES[name])
        if arg is not None:
            words.append(labels[arg] if name == "call" else int(arg) % P)
    return words
