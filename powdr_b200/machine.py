"""Host-side mirror of the reference's SymbolicMachine handling for the GPU path.

* JSON schema of `SymbolicMachine` / `AlgebraicExpression` (/root/reference/expression/src/lib.rs:209-246,
  /root/reference/autoprecompiles/src/symbolic_machine.rs:35-39,70-75,114-124): int = canonical constant,
  "name@id" = column reference, [l, op, r] with op in + - *, ["-", e] unary minus.
* column order = ascending poly id of the referenced columns (`main_columns()` -> `unique_references()`,
  /root/reference/autoprecompiles/src/powdr.rs:44-57; `apc_poly_id_to_index`, trace_handler.rs:70-75).
* bytecode emission = `emit_expr` / `compile_derived_to_gpu` / `compile_bus_to_gpu`
  (/root/reference/openvm/src/powdr_extension/trace_generator/cuda/mod.rs:49-177), post-order, same opcodes
  (/root/reference/openvm/src/cuda_abi.rs:138-147).
* the text form of the optimized-APC snapshots (/root/reference/openvm-riscv/tests/apc_snapshots/**) is parsed with the
  precedence of /root/reference/expression/src/display.rs:21-66.
"""
import gzip
import json
import re

P = 2013265921
OP_PUSH_APC, OP_PUSH_CONST, OP_ADD, OP_SUB, OP_MUL, OP_NEG, OP_INV_OR_ZERO = range(7)
_BIN = {"+": OP_ADD, "-": OP_SUB, "*": OP_MUL}


def _ref_id(s):
    return int(s.rsplit("@", 1)[1])


def iter_refs(expr):
    stack = [expr]
    while stack:
        e = stack.pop()
        if isinstance(e, str):
            yield e
        elif isinstance(e, list):
            if len(e) == 3:
                stack.append(e[0])
                stack.append(e[2])
            else:
                stack.append(e[1])


def emit_expr(bc, expr, col_of, scale=1):
    """Append post-order bytecode for `expr`; PUSH_APC operand = col_of(ref) * scale (scale = height for the reference's
    absolute-offset convention, 1 for the column-index convention of pb_air_compile)."""
    work = [(expr, False)]
    while work:
        e, done = work.pop()
        if isinstance(e, int):
            bc += [OP_PUSH_CONST, e % P]
        elif isinstance(e, str):
            bc += [OP_PUSH_APC, col_of(e) * scale]
        elif len(e) == 3:
            if done:
                bc.append(_BIN[e[1]])
            else:
                work.append((e, True))
                work.append((e[2], False))
                work.append((e[0], False))
        else:
            if done:
                bc.append(OP_NEG)
            else:
                work.append((e, True))
                work.append((e[1], False))
    return bc


def degree(expr):
    """AlgebraicExpression::degree (/root/reference/expression/src/lib.rs:111-123)."""
    out = []
    work = [(expr, False)]
    while work:
        e, done = work.pop()
        if isinstance(e, int):
            out.append(0)
        elif isinstance(e, str):
            out.append(1)
        elif len(e) == 3:
            if done:
                r, l = out.pop(), out.pop()
                out.append(l + r if e[1] == "*" else max(l, r))
            else:
                work += [(e, True), (e[2], False), (e[0], False)]
        else:
            if not done:
                work += [(e, True), (e[1], False)]
    return out[0]


class SymbolicMachine:
    def __init__(self, constraints, bus_interactions=(), derived_columns=()):
        self.constraints = list(constraints)
        self.bus_interactions = list(bus_interactions)
        self.derived_columns = list(derived_columns)
        ids = {}
        for e in self.constraints:
            for r in iter_refs(e):
                ids[_ref_id(r)] = r
        for b in self.bus_interactions:
            for e in [b["mult"]] + list(b["args"]):
                for r in iter_refs(e):
                    ids[_ref_id(r)] = r
        for name, method in self.derived_columns:
            ids[_ref_id(name)] = name
        self.column_ids = sorted(ids)                       # ascending poly id
        self.id_to_index = {pid: i for i, pid in enumerate(self.column_ids)}
        self.column_names = [ids[i] for i in self.column_ids]

    @property
    def width(self):
        return len(self.column_ids)

    def col_of(self, ref):
        return self.id_to_index[_ref_id(ref)]

    @classmethod
    def from_json_file(cls, path):
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rt") as f:
            doc = json.load(f)
        m = doc["machine"] if "machine" in doc else doc
        return cls(m["constraints"], m.get("bus_interactions", []), m.get("derived_columns", []))

    @classmethod
    def from_snapshot_text(cls, text):
        """Parse the `// Algebraic constraints:` and bus sections of an apc_snapshots/*.txt file."""
        names = []
        in_cols = False
        constraints, buses, bus_id = [], [], None
        section = None
        for line in text.splitlines():
            s = line.strip()
            if s.startswith("Symbolic machine using"):
                in_cols = True
                continue
            if in_cols:
                if not s:
                    in_cols = False
                else:
                    names.append(s)
                continue
            m = re.match(r"// Bus (\d+)", s)
            if m:
                section, bus_id = "bus", int(m.group(1))
                continue
            if s.startswith("// Algebraic constraints"):
                section = "con"
                continue
            if not s or s.startswith("//"):
                continue
            if section == "con" and s.endswith("= 0"):
                constraints.append(s[: -len("= 0")].strip())
            elif section == "bus" and s.startswith("mult="):
                mult, args = s[len("mult="):].split(", args=[", 1)
                buses.append((bus_id, mult, _split_top(args[:-1])))
        ident = {n: "%s@%d" % (n, i) for i, n in enumerate(names)}
        cons = [_parse_text_expr(c, ident) for c in constraints]
        bis = [{"id": b, "mult": _parse_text_expr(mu, ident), "args": [_parse_text_expr(a, ident) for a in ar]} for b, mu, ar in buses]
        mach = cls(cons, bis, [])
        mach.snapshot_columns = names
        return mach


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


_TOK = re.compile(r"\s*(\d+|[A-Za-z_][A-Za-z_0-9]*|[()+\-*])")


def _parse_text_expr(text, ident):
    toks = _TOK.findall(text)
    pos = [0]

    def peek():
        return toks[pos[0]] if pos[0] < len(toks) else None

    def take():
        t = toks[pos[0]]
        pos[0] += 1
        return t

    def atom():
        t = take()
        if t == "(":
            e = sum_()
            assert take() == ")"
            return e
        if t == "-":
            return ["-", prod()]
        if t.isdigit():
            return int(t) % P
        return ident[t]

    def prod():
        e = atom()
        while peek() == "*":
            take()
            e = [e, "*", atom()]
        return e

    def sum_():
        e = prod()
        while peek() in ("+", "-"):
            op = take()
            e = [e, op, prod()]
        return e

    e = sum_()
    assert pos[0] == len(toks), (text, toks[pos[0]:])
    return e


def compile_constraints(machine):
    """-> (bytecode list, [(off, len)]) in the column-index convention of pb_air_compile."""
    bc, spans = [], []
    for c in machine.constraints:
        off = len(bc)
        emit_expr(bc, c, machine.col_of, 1)
        spans.append((off, len(bc) - off))
    return bc, spans


def compile_derived(machine, height):
    """compile_derived_to_gpu (cuda/mod.rs:100-141): -> ([(apc_col_index, off, len)], bytecode) with absolute offsets."""
    specs, bc = [], []
    for name, method in machine.derived_columns:
        col = machine.col_of(name)
        off = len(bc)
        if "Constant" in method:
            bc += [OP_PUSH_CONST, int(method["Constant"]) % P]
        else:
            e1, e2 = method["QuotientOrZero"]
            emit_expr(bc, e2, machine.col_of, height)
            bc.append(OP_INV_OR_ZERO)
            emit_expr(bc, e1, machine.col_of, height)
            bc.append(OP_MUL)
        specs.append((col, off, len(bc) - off))
    return specs, bc


def compile_bus(machine, height):
    """compile_bus_to_gpu (cuda/mod.rs:143-177): -> (interactions [(bus_id, num_args, args_index_off)], arg_spans, bytecode)."""
    ints, spans, bc = [], [], []
    for b in machine.bus_interactions:
        off_idx = len(spans)
        for e in [b["mult"]] + list(b["args"]):
            off = len(bc)
            emit_expr(bc, e, machine.col_of, height)
            spans.append((off, len(bc) - off))
        ints.append((int(b["id"]), len(b["args"]), off_idx))
    return ints, spans, bc


def periphery_from_bus_map(bus_map):
    """The periphery bus ids and table sizes stage 0's histogram kernel needs (`self.periphery.bus_ids.*`, `tuple_range_checker_chip.sizes`,
    /root/reference/openvm/src/powdr_extension/trace_generator/cuda/mod.rs:359-372) out of the `bus_map` section of the reference's APC
    JSON export (`BusMap<OpenVmBusType>`): {"bus_ids": {"3": {"Other": "VariableRangeChecker"}, "6": {"Other":
    "BitwiseLookup"}, "7": {"Other": {"TupleRangeChecker": [256, 2048]}}, ...}}.
    -> dict(var_bus, bitwise_bus, tuple2_bus, tuple2_sizes); a missing periphery bus maps to None."""
    out = {"var_bus": None, "bitwise_bus": None, "tuple2_bus": None, "tuple2_sizes": None}
    for bus_id, kind in (bus_map or {}).get("bus_ids", {}).items():
        other = kind.get("Other") if isinstance(kind, dict) else None
        if other == "VariableRangeChecker":
            out["var_bus"] = int(bus_id)
        elif other == "BitwiseLookup":
            out["bitwise_bus"] = int(bus_id)
        elif isinstance(other, dict) and "TupleRangeChecker" in other:
            out["tuple2_bus"] = int(bus_id)
            out["tuple2_sizes"] = tuple(int(x) for x in other["TupleRangeChecker"])
    return out


# OpenVM RV32IM opcode classes (first global opcode, number of opcodes, executing AIR) -- RECOLLECTED from the `#[opcode_offset = ..]`
# attributes of openvm's rv32im transpiler crate (not in the reference tree).  What the tree does confirm: in the keccak fixture the
# opcodes of one class agree on the original AIR's width (0x200-0x204: 36 columns, 0x205/0x206: 53, 0x210/0x213: 41, 0x221: 26,
# 0x231: 18) and `single_div_nondet` is opcode 0x254 (tests/test_oracle.py::test_rv32_opcode_classes_agree_with_the_fixture_widths).
RV32_OPCODE_CLASSES = [
    (0x200, 5, "BaseAlu"), (0x205, 3, "Shift"), (0x208, 2, "LessThan"), (0x210, 6, "LoadStore"), (0x216, 2, "LoadSignExtend"),
    (0x220, 2, "BranchEqual"), (0x225, 4, "BranchLessThan"), (0x230, 2, "JalLui"), (0x235, 1, "Jalr"), (0x240, 1, "Auipc"),
    (0x250, 1, "Mul"), (0x251, 3, "MulH"), (0x254, 4, "DivRem"),
]


def rv32_air_of_opcode(op):
    """opcode -> AIR key by opcode class (the role of `original_airs.opcode_to_air`); unknown opcodes are their own AIR"""
    for first, count, name in RV32_OPCODE_CLASSES:
        if first <= op < first + count:
            return name
    return op


def compile_substitutions(opcodes, subs, id_to_index, opcode_to_air=None):
    """The (`OriginalAir`, `Subst`) tables of stage 0 from an APC's instruction list and its per-instruction substitutions, the way
    `PowdrTraceGeneratorGpu::try_generate_witness` assembles them (/root/reference/openvm/src/powdr_extension/trace_generator/cuda/mod.rs:268-326):
    instructions with substitutions are grouped by the original AIR their opcode executes on; an AIR's group index is its
    `air_index`, an instruction's position inside its group is the `row` of its cells in that AIR's block of rows per APC call
    (`row_block_size` = group size), `col` = `original_poly_index`, `apc_col` = index of `apc_poly_id` among the machine's columns.

    opcodes: one opcode per instruction (`block.blocks[*].instructions[*][0]` of the reference's APC JSON);
    subs: per instruction [(original_poly_index, apc_poly_id)]; id_to_index: apc poly id -> column index (SymbolicMachine.id_to_index);
    opcode_to_air: opcode -> AIR key (`original_airs.opcode_to_air`, an OpenVM table that is not in the reference tree; default: one
    AIR per opcode).  Rust groups with a HashMap, so the AIR order is arbitrary there; here it is the order of first appearance.
    -> (airs [(air_key, row_block_size, min_width)], substs [(air_index, col, row, apc_col)])"""
    assert len(opcodes) == len(subs)
    air_of = opcode_to_air or (lambda op: op)
    groups, order = {}, []
    for op, ss in zip(opcodes, subs):
        if not ss:
            continue                                    # an instruction without substitutions contributes no rows (cuda/mod.rs:277-281)
        key = air_of(op)
        if key not in groups:
            groups[key] = []
            order.append(key)
        groups[key].append(ss)
    airs, substs = [], []
    for air_index, key in enumerate(order):
        rows = groups[key]
        width = 0
        for row, ss in enumerate(rows):
            for orig, apc_id in ss:
                substs.append((air_index, int(orig), row, id_to_index[int(apc_id)]))
                width = max(width, int(orig) + 1)
        airs.append((key, len(rows), width))
    return airs, substs


def synthetic_machine(width, n_constraints, seed=0):
    """Synthetic AIR with the pinned post-optimisation SHAPE of an APC (width, constraint count, degree <= 3;
    SURVEY.md §8d): column i belongs to constraint i % C; a constraint is the sum, over consecutive triples (x, y, z) of its
    columns, of  x*y*z,  x*(y + c*z)  or  c*x - y*z  -- every column referenced, no constant term, so it vanishes on
    the all-zero (padding) row like a guarded APC (/root/reference/autoprecompiles/src/lib.rs:415-453)."""
    s = [(seed * 0x9E3779B97F4A7C15 + 0xB2000000) & (2**64 - 1)]

    def rnd():
        s[0] = (s[0] + 0x9E3779B97F4A7C15) & (2**64 - 1)
        z = s[0]
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        return z ^ (z >> 31)

    n_constraints = min(n_constraints, width)
    names = ["c%d@%d" % (i, i) for i in range(width)]
    cons = []
    for k in range(n_constraints):
        cols = names[k::n_constraints]
        acc = None
        for i in range(0, len(cols), 3):
            tri = cols[i:i + 3]
            c = 2 + rnd() % (P - 2)
            kind = rnd() % 3
            if len(tri) == 3:
                x, y, z = tri
                term = [[x, "*", y], "*", z] if kind == 0 else ([x, "*", [y, "+", [c, "*", z]]] if kind == 1 else [[c, "*", x], "-", [y, "*", z]])
            elif len(tri) == 2:
                term = [tri[0], "*", tri[1]] if kind else [tri[0], "-", [c, "*", tri[1]]]
            else:
                term = [c, "*", tri[0]] if kind else tri[0]
            acc = term if acc is None else [acc, "+", term]
        cons.append(acc)
    return SymbolicMachine(cons)


def synthetic_bus(machine_or_width, n_interactions, seed=0, quadratic_every=0):
    """Synthetic bus interactions with the SHAPE of an APC's: the bus mix of the reference's keccak fixture
    (/root/reference/autoprecompiles/tests/keccak_apc_pre_opt.json.gz, SURVEY.md App. A: bus 3 variable range checker 39 %,
    bus 1 memory 31 %, bus 6 bitwise lookup 16 %, bus 0 execution bridge 10 %, bus 2 pc lookup 5 %) and their argument
    tuples (periphery.rs:179-236: bitwise [x, y, x^y, selector], range checker [value, max_bits]; memory
    [address space, pointer, 4 data limbs, timestamp]; execution bridge [pc, timestamp]; pc lookup 9 words), every
    expression of degree <= 1 (mult is a column or a negated column -- a guarded APC multiplies by is_valid), so two
    interactions share a LogUp chunk.  quadratic_every > 0 makes every such interaction carry one degree-2 argument (bus
    degree bound 2, openvm/src/lib.rs:97-101), which forces a chunk of its own.
    Returns the list in the JSON schema of SymbolicBusInteraction ({"id", "mult", "args"})."""
    width = machine_or_width if isinstance(machine_or_width, int) else machine_or_width.width
    names = ["c%d@%d" % (i, i) for i in range(width)] if isinstance(machine_or_width, int) else machine_or_width.column_names
    s = [(seed * 0x9E3779B97F4A7C15 + 0xB05B05) & (2**64 - 1)]

    def rnd():
        s[0] = (s[0] + 0x9E3779B97F4A7C15) & (2**64 - 1)
        z = s[0]
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        return z ^ (z >> 31)

    def col():
        return names[rnd() % width]

    def lin():
        k = rnd() % 4
        if k == 0:
            return [col(), "+", int(rnd() % 65536)]
        if k == 1:
            return [col(), "-", [int(2 + rnd() % (P - 2)), "*", col()]]
        return col()

    out = []
    for i in range(n_interactions):
        t = rnd() % 100
        mult = col() if rnd() % 2 else ["-", col()]
        if t < 39:
            bus, args = 3, [lin(), int(1 + rnd() % 17)]
        elif t < 70:
            bus, args = 1, [int(1 + rnd() % 2), lin(), col(), col(), col(), col(), lin()]
        elif t < 86:
            bus, args = 6, [col(), col(), col(), int(rnd() % 2)]
        elif t < 96:
            bus, args = 0, [int(rnd() % (1 << 22)), lin()]
        else:
            bus, args = 2, [int(rnd() % (1 << 22))] + [int(rnd() % 4096) for _ in range(8)]
        if quadratic_every and i % quadratic_every == quadratic_every - 1:
            args[0] = [col(), "*", col()]
        out.append({"id": bus, "mult": mult, "args": args})
    return out
