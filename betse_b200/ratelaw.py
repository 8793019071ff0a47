"""Rate-law compiler: the reference's generated expression strings -> bytecode for the GPU.

``MasterOfNetworks`` (betse/science/chemistry/networks.py) writes every rate law of a general
network / gene regulatory network as a *Python expression string* and ``eval``s it each timestep:

* growth/decay of a substance — ``write_growth_and_decay`` (networks.py:1187-1310):
  ``mod * r_prod * alpha_growth - r_decay * c - decay_max * alpha_decay * c``;
* cell-zone reactions — ``write_reactions`` (networks.py:1312-1572);
* channel modulation — ``alpha_eval_string`` of a ``Channel`` (networks.py:1102-1110, 3147);
  all of them built from the Hill algebra of ``get_influencers`` (networks.py:5084-5602).

Instead of re-deriving that algebra from the YAML (and every ``!`` / ``&`` / ``*`` suffix rule with
it), this module takes the strings the reference itself produced, parses them with ``ast`` and
emits a postfix program over

* *dynamic* leaves, which live on the device: ``self.cell_concs['X']`` /
  ``self.molecules['X'].c_cells`` (substance X in the cell), ``self.mem_concs['X']`` (X at the
  membrane = X of the membrane's cell when intracellular mixing is instant, networks.py:5722-5724),
  ion names in the same dictionaries, ``sim.vm``;
* *static* leaves, resolved ONCE on the host by a ``resolver(source) -> value`` callback (the live
  reference objects in the drop-in loop; the recorded values of a golden fixture in the tests) and
  constant-folded with the same IEEE operations Python would have applied.

Operations keep Python's evaluation order, so the device result differs from the reference's
``eval`` only by ``pow``/``exp`` rounding.  Anything outside this vocabulary raises
``RateLawError`` — the drop-in loop then refuses the configuration instead of approximating it.
"""
import ast
import operator

import numpy as np

# opcodes (mirrored in csrc/network.cuh)
PUSHC, PUSHS, PUSHA, PUSHI, PUSHM, PUSHV, PUSHE, PUSHJ, ADD, SUB, MUL, DIV, POW, NEG, EXP = range(15)
OP_NAMES = ["PUSHC", "PUSHS", "PUSHA", "PUSHI", "PUSHM", "PUSHV", "PUSHE", "PUSHJ", "ADD", "SUB", "MUL", "DIV", "POW", "NEG", "EXP"]
LAST_PUSH = PUSHJ
MAX_STACK = 16


class RateLawError(Exception):
    pass


_BIN = {ast.Add: (ADD, operator.add), ast.Sub: (SUB, operator.sub), ast.Mult: (MUL, operator.mul),
        ast.Div: (DIV, operator.truediv), ast.Pow: (POW, operator.pow)}


class Program:
    """Postfix code + the tables it indexes.  One ``Tables`` object is shared by all programs of a
    network handler so that constants and constant arrays are uploaded once."""

    def __init__(self, code, zone):
        self.code = code          # list of (op, arg)
        self.zone = zone          # 'cell' or 'mem'

    def max_depth(self):
        d = m = 0
        for op, _ in self.code:
            if op <= LAST_PUSH:
                d += 1
            elif op in (NEG, EXP):
                pass
            else:
                d -= 1
            m = max(m, d)
        return m


class Tables:
    def __init__(self, species, ions, n_cells, n_mems, n_env=0):
        self.species = list(species)
        self.ions = list(ions)
        self.n_cells, self.n_mems, self.n_env = int(n_cells), int(n_mems), int(n_env)
        self.consts = []
        self.cell_arrays = []
        self.mem_arrays = []
        self.env_species = set()      # substances some program reads outside the membrane (need env storage)

    def const(self, v):
        v = float(v)
        for k, c in enumerate(self.consts):
            if c == v and np.signbit(c) == np.signbit(v):
                return k
        self.consts.append(v)
        return len(self.consts) - 1

    def array(self, a, zone):
        if zone == "env":
            raise RateLawError("a non-uniform static array in the extracellular zone is not implemented")
        tab = self.cell_arrays if zone == "cell" else self.mem_arrays
        for k, b in enumerate(tab):
            if np.array_equal(a, b):
                return k
        tab.append(np.ascontiguousarray(a, dtype=np.float64))
        return len(tab) - 1


def _subscript_key(node):
    s = node.slice
    if isinstance(s, ast.Constant) and isinstance(s.value, str):
        return s.value
    return None


def _dynamic(node, tables, zone):
    """(op, arg) if ``node`` is a device-resident quantity, else None."""
    # self.env_concs['X'][cells.map_mem2ecm]: X outside the membrane (transporters, networks.py:2346-2430)
    if isinstance(node, ast.Subscript) and isinstance(node.value, ast.Subscript) \
            and isinstance(node.value.value, ast.Attribute) and node.value.value.attr == "env_concs" \
            and isinstance(node.value.value.value, ast.Name) and node.value.value.value.id == "self":
        key = _subscript_key(node.value)
        idx = ast.unparse(node.slice)
        # membrane zone: the env square of the membrane; cell zone: the env square of the cell centre (get_influencers,
        # networks.py:5242-5265)
        if (zone, idx) not in (("mem", "cells.map_mem2ecm"), ("cell", "cells.map_cell2ecm")):
            raise RateLawError("env_concs[%r][%s] in the %s zone is not implemented" % (key, idx, zone))
        if key in tables.species:
            tables.env_species.add(tables.species.index(key))
            return (PUSHE, tables.species.index(key))
        if key in tables.ions:
            return (PUSHJ, tables.ions.index(key))
        raise RateLawError("unknown substance %r" % key)
    # self.cell_concs['X'] / self.mem_concs['X'] / self.env_concs[...]
    if isinstance(node, ast.Subscript) and isinstance(node.value, ast.Attribute) \
            and isinstance(node.value.value, ast.Name) and node.value.value.id == "self":
        dic, key = node.value.attr, _subscript_key(node)
        if dic == "env_concs" and zone == "env" and key is not None:
            # extracellular zone (tight-junction modulators, get_influencers with reaction_zone 'env', networks.py:5270-5281):
            # the whole env array, evaluated square by square
            if key in tables.species:
                tables.env_species.add(tables.species.index(key))
                return (PUSHE, tables.species.index(key))
            if key in tables.ions:
                return (PUSHJ, tables.ions.index(key))
            raise RateLawError("unknown substance %r" % key)
        if dic in ("cell_concs", "mem_concs") and key is not None:
            if (dic == "cell_concs") != (zone == "cell"):
                raise RateLawError("%s[%r] used in the %s zone" % (dic, key, zone))
            if key in tables.species:
                return (PUSHS, tables.species.index(key))
            if key in tables.ions:
                # cell zone: sim.cc_cells[ion] (networks.py:186-206).  Membrane zone: sim.cc_at_mem[ion], which during the
                # network block of a step IS sim.cc_cells[ion][mem_to_cells] — set so at the end of the ion loop
                # (update_intra, sim.py:2310) and by every update_Co since (sim_toolbox.py:1182) — so the same read
                # serves both zones (a transporter's own tweak of mem_concs, networks.py:3020-3022, lives only until
                # the next update_Co and is not reproduced)
                return (PUSHI, tables.ions.index(key))
            raise RateLawError("unknown substance %r" % key)
        if dic in ("env_concs", "mit_concs", "bound_concs"):
            raise RateLawError("%s[...] (extracellular / mitochondrial zone) is not implemented" % dic)
    # self.molecules['X'].c_cells
    if isinstance(node, ast.Attribute) and node.attr in ("c_cells", "cc_at_mem", "c_env") \
            and isinstance(node.value, ast.Subscript) and isinstance(node.value.value, ast.Attribute) \
            and node.value.value.attr == "molecules":
        key = _subscript_key(node.value)
        if node.attr == "c_cells" and zone == "cell" and key in tables.species:
            return (PUSHS, tables.species.index(key))
        if node.attr == "cc_at_mem" and zone == "mem" and key in tables.species:
            return (PUSHS, tables.species.index(key))
        raise RateLawError("molecules[%r].%s in the %s zone is not implemented" % (key, node.attr, zone))
    # sim.vm
    if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == "sim" \
            and node.attr == "vm":
        if zone != "mem":
            raise RateLawError("sim.vm outside the membrane zone")
        return (PUSHV, 0)
    return None


def _has_dynamic(node, tables, zone):
    for n in ast.walk(node):
        try:
            if _dynamic(n, tables, zone) is not None:
                return True
        except RateLawError:
            return True
    return False


def compile_expr(src, tables, resolver, zone="cell"):
    """Compile one expression string.  ``resolver(source_text)`` evaluates a static sub-expression
    (an attribute chain such as ``self.molecules['X'].r_production`` or a call such as
    ``np.ones(sim.cdl)``) and returns a float or an array of the zone's length."""
    tree = ast.parse(src.strip(), mode="eval").body
    n_zone = tables.n_cells if zone == "cell" else (tables.n_env if zone == "env" else tables.n_mems)

    def static_value(v, what):
        if isinstance(v, (bool, np.bool_)):
            v = float(v)
        if isinstance(v, (int, float, np.integer, np.floating)):
            return v
        a = np.asarray(v)
        if a.ndim == 0:
            return float(a)
        if a.ndim == 1 and a.shape[0] == n_zone and a.dtype.kind in "fiub":
            return a.astype(np.float64) if a.dtype.kind != "f" else a
        raise RateLawError("static term %r has shape %s, expected a scalar or [%d]" % (what, a.shape, n_zone))

    def push_static(v):
        if isinstance(v, np.ndarray):
            if v.size and np.all(v == v[0]) and not np.signbit(v[0]):
                return [(PUSHC, tables.const(float(v[0])))]      # np.ones(n) and friends
            return [(PUSHA, tables.array(v, zone))]
        return [(PUSHC, tables.const(float(v)))]

    def ev(node):
        """-> ('s', value) for a host-side constant, ('d', code) for device code."""
        if isinstance(node, ast.Constant):
            if isinstance(node.value, (int, float)) and not isinstance(node.value, bool):
                return ("s", node.value)
            raise RateLawError("unsupported literal %r" % (node.value,))
        dyn = _dynamic(node, tables, zone)
        if dyn is not None:
            return ("d", [dyn])
        if isinstance(node, ast.BinOp) and type(node.op) in _BIN:
            op, fn = _BIN[type(node.op)]
            a, b = ev(node.left), ev(node.right)
            if a[0] == "s" and b[0] == "s":
                with np.errstate(all="ignore"):
                    return ("s", static_value(fn(a[1], b[1]), ast.unparse(node)))
            ca = a[1] if a[0] == "d" else push_static(a[1])
            cb = b[1] if b[0] == "d" else push_static(b[1])
            return ("d", ca + cb + [(op, 0)])
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            a = ev(node.operand)
            if isinstance(node.op, ast.UAdd):
                return a
            if a[0] == "s":
                return ("s", -a[1])
            return ("d", a[1] + [(NEG, 0)])
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) \
                and isinstance(node.func.value, ast.Name) and node.func.value.id == "np" \
                and node.func.attr == "exp" and len(node.args) == 1 and _has_dynamic(node.args[0], tables, zone):
            a = ev(node.args[0])
            return ("d", a[1] + [(EXP, 0)])
        if _has_dynamic(node, tables, zone):
            for sub in ast.walk(node):                 # a more specific reason, if one of the leaves has it
                _dynamic(sub, tables, zone)
            raise RateLawError("unsupported construct around a concentration: %s" % ast.unparse(node))
        # a static leaf (attribute chain, call, name): ask the host
        text = ast.unparse(node)
        return ("s", static_value(resolver(text), text))

    kind, val = ev(tree)
    code = val if kind == "d" else push_static(val)
    prog = Program(code, zone)
    if prog.max_depth() > MAX_STACK:
        raise RateLawError("expression needs a stack deeper than %d" % MAX_STACK)
    return prog


def live_resolver(core, sim, p, cells, record=None):
    """Resolver over the live reference objects (the namespace networks.py evals in)."""
    ns = {"self": core, "sim": sim, "p": p, "cells": cells, "np": np}

    def resolve(text):
        v = eval(text, {"np": np}, ns)   # noqa: S307 - reference-generated attribute chains only
        if record is not None:
            record[text] = np.asarray(v, dtype=np.float64) if np.ndim(v) else float(v)
        return v
    return resolve


def table_resolver(table):
    """Resolver over recorded values (golden fixtures: ``{source text: value}``)."""
    def resolve(text):
        if text not in table:
            raise RateLawError("no recorded value for static term %r" % text)
        v = table[text]
        return float(v) if np.ndim(v) == 0 else np.asarray(v)
    return resolve


def pack_programs(programs):
    """-> (int32 code [2*len], int32 ptr [n+1]) for the C ABI."""
    code, ptr = [], [0]
    for pr in programs:
        for op, arg in pr.code:
            code += [int(op), int(arg)]
        ptr.append(len(code) // 2)
    return np.asarray(code if code else [0, 0], dtype=np.int32), np.asarray(ptr, dtype=np.int32)


def run_numpy(prog, tables, species, ions=None, ions_mid=None, vm=None, mem_to_cells=None, species_env=None, ions_env=None,
              map_mem2ecm=None, map_cell2ecm=None):
    """Host interpreter of a program (tests only): ``species`` [K][C]; membrane-zone programs
    gather cell quantities through ``mem_to_cells``."""
    st = []
    g = (lambda a: a[mem_to_cells]) if prog.zone == "mem" else (lambda a: a)
    env_ix = slice(None) if prog.zone == "env" else (map_mem2ecm if prog.zone == "mem" else map_cell2ecm)
    for op, arg in prog.code:
        if op == PUSHC:
            st.append(tables.consts[arg])
        elif op == PUSHS:
            st.append(g(species[arg]))
        elif op == PUSHA:
            st.append((tables.cell_arrays if prog.zone == "cell" else tables.mem_arrays)[arg])
        elif op == PUSHI:
            st.append(g(ions[arg]))
        elif op == PUSHM:
            st.append(g(ions_mid[arg]))
        elif op == PUSHV:
            st.append(vm)
        elif op == PUSHE:
            st.append(species_env[arg][env_ix])
        elif op == PUSHJ:
            st.append(ions_env[arg][env_ix])
        elif op == NEG:
            st.append(-st.pop())
        elif op == EXP:
            st.append(np.exp(st.pop()))
        else:
            b = st.pop()
            a = st.pop()
            with np.errstate(all="ignore"):
                st.append({ADD: operator.add, SUB: operator.sub, MUL: operator.mul, DIV: operator.truediv,
                           POW: np.power}[op](a, b))
    assert len(st) == 1
    n = tables.n_cells if prog.zone == "cell" else (tables.n_env if prog.zone == "env" else tables.n_mems)
    return st[0] * np.ones(n)
