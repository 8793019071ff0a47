"""Host logic of the network path: the rate-law compiler (betse_b200/ratelaw.py)
against Python ``eval`` of the same reference-generated strings, and the network description round trip."""
import numpy as np
import pytest

from betse_b200 import network as netlib
from betse_b200 import ratelaw
from oracle.betse_oracle import OracleSim
from tests import util


def _setup(kind="sim"):
    cap = util.load_golden("mammal_ecm_net")
    descs = util.networks_of(cap, kind)
    o = OracleSim(util.group(cap, "cells."), util.group(cap, kind + ".p."), util.group(cap, kind + ".s0."),
                  channels=util.channels_of(cap, kind), phase_init=(kind == "init"), networks=descs)
    return cap, descs, o


def test_compiled_programs_equal_python_eval():
    """Every growth/decay, reaction and channel-modulator string: bytecode (run on the host interpreter)
    == eval of the string, on random concentrations."""
    cap, descs, o = _setup()
    d, net = descs[0], o.networks[0]
    C, M = o.cdl, o.mdl
    comp = netlib.compile_network(d, C, M)
    rng = np.random.default_rng(7)
    for trial in range(3):
        for n in net.species:
            net.c[n] = rng.uniform(0.0, 3.0, C)
            net.cmem[n] = net.c[n][o.mem_to_cells]
        spec = np.stack([net.c[n] for n in net.species])
        strings = d["gad_strings"] + d["reaction_strings"]
        for pr, s in zip(comp["rate_programs"], strings):
            want = net.eval_string(s) * np.ones(C)
            got = ratelaw.run_numpy(pr, comp["tables"], spec, ions=o.cc_cells)
            assert np.allclose(got, want, rtol=1e-14, atol=0), s
        mods = iter(comp["mod_programs"])
        for s, idx in zip(d["chan_mod_strings"], comp["mod_index"]):
            want = net.eval_string(s) * np.ones(M)
            if idx < 0:
                assert np.all(want == 1.0)
                continue
            got = ratelaw.run_numpy(next(mods), comp["tables"], spec, vm=o.vm, mem_to_cells=o.mem_to_cells)
            assert np.allclose(got, want, rtol=1e-14, atol=0), s
    assert comp["mod_index"] == [-1, 5, 6, -1]          # Nav and Cav are not modulated
    assert comp["growth_mask"] is not None and comp["growth_mask"][3].sum() == 4   # X grows in the 'Spot' profile only


def test_static_terms_are_folded():
    cap, descs, o = _setup()
    comp = netlib.compile_network(descs[0], o.cdl, o.mdl)
    for pr in comp["rate_programs"]:
        ops = [op for op, _ in pr.code]
        assert ratelaw.PUSHA not in ops                   # np.ones(...) and uniform profiles fold to constants
        assert pr.max_depth() <= ratelaw.MAX_STACK


@pytest.mark.parametrize("src,msg", [
    ("self.env_concs['X'][cells.map_mem2ecm]/2.0", "cell zone is not implemented"),      # the membrane map in a cell-zone law
    ("self.env_concs['X']/2.0", "extracellular"),                                       # a whole env field (env-zone reactions)
    ("self.cell_concs['Nope']*2", "unknown substance"),
    ("np.sin(self.cell_concs['X'])", "unsupported construct"),
    ("self.mem_concs['X']*2", "zone"),
])
def test_unsupported_expressions_are_refused(src, msg):
    tabs = ratelaw.Tables(["X"], ["Na", "K"], 4, 12)
    with pytest.raises(ratelaw.RateLawError) as e:
        ratelaw.compile_expr(src, tabs, ratelaw.table_resolver({}), "cell")
    assert msg in str(e.value)


def test_description_roundtrip():
    cap, descs, _ = _setup()
    flat = netlib.flatten(descs[0], "x.")
    back = netlib.unflatten(flat, "x.")
    assert back["gad_strings"] == descs[0]["gad_strings"] and back["species"] == descs[0]["species"]
    assert set(back["static"]) == set(descs[0]["static"])
    for k, v in back["static"].items():
        assert np.array_equal(np.asarray(v), np.asarray(descs[0]["static"][k]))


def test_transporter_programs_equal_python_eval():
    """Transporter rate laws (write_transporters, networks.py:2090-2614): concentrations outside the membrane
    (env_concs[...][cells.map_mem2ecm]), ions at the membrane, np.exp of the electrochemical term and Vmem."""
    cap = util.load_golden("mammal_ecm_net_trans")
    kind = "sim"
    descs = util.networks_of(cap, kind)
    o = OracleSim(util.mesh_of(cap, kind), util.group(cap, kind + ".p."), util.group(cap, kind + ".s0."), networks=descs)
    d, net = descs[0], o.networks[0]
    comp = netlib.compile_network(d, o.cdl, o.mdl)
    assert len(comp["transporters"]) == 2 and comp["transporters"][1]["net_z"] == 3.0
    kinds = sorted((k, sg) for k, _, sg, _ in comp["transporters"][1]["terms"])
    assert kinds == [(0, -1), (1, 1), (2, -1), (2, 1)]          # Na leaves the cells for the env grid; T2 -> G1 inside
    rng = np.random.default_rng(11)
    for n in net.species:
        net.c[n] = rng.uniform(0.05, 2.0, o.cdl)
        net.cmem[n] = net.c[n][o.mem_to_cells]
        if n in net.c_env:
            net.c_env[n] = rng.uniform(0.05, 2.0, o.n_env)
    o.vm = rng.uniform(-0.07, 0.01, o.mdl)
    o.cc_at_mem = o.cc_cells[:, o.mem_to_cells].copy()        # as the ion loop leaves it before the network block (sim.py:2310)
    spec = np.stack([net.c[n] for n in net.species])
    spec_env = np.stack([net.c_env.get(n, np.zeros(o.n_env)) for n in net.species])
    for t, ct in zip(d["transporters"], comp["transporters"]):
        want = net.eval_string(t["eval_string"]) * np.ones(o.mdl)
        pr = (comp["rate_programs"] + comp["mod_programs"])[ct["prog"]]
        got = ratelaw.run_numpy(pr, comp["tables"], spec, ions=o.cc_cells, vm=o.vm, mem_to_cells=o.mem_to_cells,
                                species_env=spec_env, ions_env=o.cc_env, map_mem2ecm=o.map_mem2ecm)
        # (1 - Q/Keq) cancels when the transporter is near equilibrium: judge against the size of the uncancelled rate
        scale = float(np.max(np.abs(want)))
        print(t["name"], float(np.max(np.abs(got - want))) / scale)
        assert np.max(np.abs(got - want)) <= 1e-12 * scale, t["name"]


@pytest.mark.parametrize("fixture", ["mammal_ecm_net_tj", "mammal_ecm_net_envrx"])
def test_extracellular_zone_programs_equal_python_eval(fixture):
    """Rate laws of the extracellular zone (get_influencers / write_reactions_env with reaction_zone 'env',
    networks.py:1830-2088, 5270-5281): tight-junction modulators and reactions in the bath read ``self.env_concs['X']`` over
    the whole grid — bytecode on the host interpreter == eval of the reference's string, on random env concentrations."""
    cap = util.load_golden(fixture)
    kind = "sim"
    descs = util.networks_of(cap, kind)
    o = OracleSim(util.mesh_of(cap, kind), util.group(cap, kind + ".p."), util.group(cap, kind + ".s0."), networks=descs)
    d, net = descs[0], o.networks[0]
    comp = netlib.compile_network(d, o.cdl, o.mdl)
    progs = comp["rate_programs"] + comp["mod_programs"]
    strings = [(s, progs[i]) for s, (t, i, _, _) in zip(d.get("modulator_strings", []), comp["modulators"]) if t == 2]
    strings += [(s, progs[i]) for s, i in zip(d.get("reaction_env_strings", []), comp["env_rx_index"])]
    assert strings and all(pr.zone == "env" for _, pr in strings)
    rng = np.random.default_rng(5)
    for trial in range(3):
        for n in net.species:
            if n in net.c_env:
                net.c_env[n] = rng.uniform(0.0, 2.0, o.n_env)
        spec = np.stack([net.c[n] for n in net.species])
        spec_env = np.stack([net.c_env.get(n, np.zeros(o.n_env)) for n in net.species])
        for s, pr in strings:
            want = net.eval_string(s) * np.ones(o.n_env)
            got = ratelaw.run_numpy(pr, comp["tables"], spec, ions=o.cc_cells, species_env=spec_env, ions_env=o.cc_env)
            assert got.shape == (o.n_env,)
            assert np.max(np.abs(got - want)) <= 1e-14 * max(float(np.max(np.abs(want))), 1e-300), s
    if fixture == "mammal_ecm_net_tj":
        assert [(t, ion) for t, _, _, ion in comp["modulators"]] == [(2, -1), (2, 0)]      # all ions; Na only
        assert comp["tj_targets"].size and comp["tj_targets"].max() < o.n_env
    else:
        assert comp["stoich_env"].tolist() == [[-1.0], [1.0], [0.0]]                       # S1 -> S2 out there


def test_cell_zone_law_reading_a_substance_an_env_reaction_moves_is_refused():
    """The device evaluates the cell-zone rate laws after the extracellular reactions of the step, the reference before
    (networks.py:2826-2889): a network that couples the two through a substance in the bath is refused, not approximated."""
    import copy
    cap = util.load_golden("mammal_ecm_net_envrx")
    d = copy.deepcopy(util.networks_of(cap, "sim")[0])
    C, M = len(cap["cells.cell_vol"]), len(cap["cells.mem_sa"])
    d["gad_strings"][2] = d["gad_strings"][2] + "*(self.env_concs['S2'][cells.map_cell2ecm]/(1 + self.env_concs['S2'][cells.map_cell2ecm]))"
    with pytest.raises(netlib.BetseB200Error, match="extracellular reaction"):
        netlib.compile_network(d, C, M)
