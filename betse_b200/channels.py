"""Voltage-gated channel library as DATA: every Hodgkin-Huxley style model of the
reference's ``betse/science/channels/{vg_na,vg_k,vg_ca,vg_cl}.py`` restated as a small table of
gating terms, so that ONE device kernel (csrc/channels.cu: k_chan) evaluates any of them.

A model gives, as functions of ``U = 1000*vm + shift`` [mV], the four quantities of
``ChannelsABC.update_mh`` (channels/channelsabc.py:40-60): mInf, mTau, hInf, hTau.  Each quantity is

* ``("T", a)``       the term a,
* ``("R", a, b)``    a/(a+b)     (steady state from opening/closing rates),
* ``("S", a, b)``    1/(a+b)     (time constant from the rates),

and a term is ``(type, p0, p1, p2, p3)``:

====  =======================================  ==============================================
type  value                                    reference form
====  =======================================  ==============================================
0     p0                                       constants (leak channels: 1)
1     p0 + p1/(1 + exp((U - p2)/p3))           Boltzmann steady states and bell-less taus
2     p0*U + p1                                linear taus (vg_k.py Kv1p3, Kv1p5)
3     p0 + p1*exp((p2 - U)/p3)                 exponential taus / rates
4     p0 + p1*exp(-((U + p2)/p3)**2)           Gaussian taus (vg_k.py K_Fast)
5     p0*(U - p1)/(1 - exp(-(U - p1)/p2))      "linoid" opening rate
6     p0*(W - p1)/(1 - exp(-(W - p1)/p2)),     "linoid" closing rate written in -V
      W = -U
7     1 where U >= p2, else                    exponential tau overwritten with 1 above a voltage
      p0 + p1*exp(-U/p3)                       (vg_ca.py Cav3p1: ``_mTau[V >= -10e-3] = 1.0``)
8     p0/cosh((U - p1)/p2)                     Morris-Lecar time constants, Kir_ML steady state
9     p0 + p1*tanh((U - p2)/p3)                Morris-Lecar steady states 0.5*(1 + tanh(...))
====  =======================================  ==============================================

The Morris-Lecar family (``vg_morrislecar.py``, selected with ``channel class: ML``, networks.py:6611-6613) has one gate and
its own update (``ChannelsABC.update_ml``, channelsabc.py:60-71): ``m = (m + dt*Phi*mInf/mTau)/(1 + dt*Phi/mTau)`` for a
kinetic gate, ``m = mInf`` otherwise; P = m.  A model carries ``ml = {"phi": Phi, "kinetic": bool}``; the device
runs it through the Hodgkin-Huxley update with the effective time constant ``mTau/Phi`` (the same quotient, multiplied
through by mTau/Phi) or 0 (``device_quantities``).

``tests/test_channels_table.py`` holds every entry to the reference's own class over a voltage
sweep (build container only: needs /root/reference).  The NumPy evaluator below is set-up / test
code (initial gate values of synthetic tissues); the per-timestep evaluation is the CUDA kernel.
"""
import numpy as np

CONST, SIG, LIN, EXP, GAUSS, LINOID, LINOID_NEG, EXP_CUT, SECH, TANH = range(10)
KIND = {"T": 0, "R": 1, "S": 2}


def _c(v):
    return (CONST, float(v), 0.0, 0.0, 0.0)


def _sig(v0, s, A=1.0, c0=0.0):
    return (SIG, float(c0), float(A), float(v0), float(s))


def _lin(a, b):
    return (LIN, float(a), float(b), 0.0, 0.0)


def _exp(A, v0, s, c0=0.0):
    """c0 + A*exp((v0 - U)/s)"""
    return (EXP, float(c0), float(A), float(v0), float(s))


def _gauss(c0, A, v0, s):
    return (GAUSS, float(c0), float(A), float(v0), float(s))


def _linoid(a, v0, s):
    return (LINOID, float(a), float(v0), float(s), 0.0)


def _linoid_neg(a, v0, s):
    return (LINOID_NEG, float(a), float(v0), float(s), 0.0)


def _T(a):
    return ("T", a)


def _leak(ion):
    one = _T(_c(1.0))
    return dict(ion=ion, time_unit=1.0, mpow=0, hpow=0, shift=0.0, q=[one, one, one, one])


def _hh(ion, mpow, hpow, mInf, mTau, hInf, hTau, shift=0.0, time_unit=1.0e3):
    return dict(ion=ion, time_unit=time_unit, mpow=mpow, hpow=hpow, shift=shift, q=[mInf, mTau, hInf, hTau])


def _multi(model, ions, rel_perm, module, init_shift=None):
    """A family that conducts several ions through one gate (vg_funny.py:73-74, cation.py:64-65): the ions are
    applied one after the other with ``DChan = P*rel_perm[j]*maxDm*moddy`` (networks.py:3158-3203)."""
    model = dict(model, ion=ions[0], ions=list(ions), rel_perm=[float(x) for x in rel_perm], module=module)
    if init_shift is not None:
        model["init_shift"] = float(init_shift)     # the class initialises its gates with another offset than it runs with
    return model


def _hcn(v0, s, tau, shift, pm_na=0.2, init_shift=None):
    m = _hh("Na", 1, 0, _T(_sig(v0, s)), _T(_c(tau)), _T(_c(1.0)), _T(_c(1.0)), shift=shift)
    return _multi(m, ["Na", "K", "Ca"], [pm_na, 1.0, 0.05], "vg_funny", init_shift)


def _ml(ions, rel_perm, mInf, mTau, phi, kinetic):
    """One Morris-Lecar model (vg_morrislecar.py): V = 1000*vm, P = m."""
    one = _T(_c(1.0))
    m = dict(ion=ions[0], time_unit=1.0e3, mpow=1, hpow=0, shift=0.0, q=[_T(mInf), _T(mTau), one, one],
             ml={"phi": float(phi), "kinetic": bool(kinetic)}, module="vg_morrislecar")
    if len(ions) > 1:
        m.update(ions=list(ions), rel_perm=[float(x) for x in rel_perm])
    return m


def _mlk(v0, s, phi=0.066):
    """Kinetic K+ channel: mInf = 0.5*(1 + tanh((V - v0)/s)), mTau = 1/cosh((V - v0)/(2 s))."""
    return _ml(["K"], [1], (TANH, 0.5, 0.5, float(v0), float(s)), (SECH, 1.0, float(v0), 2.0 * s, 0.0), phi, True)


def _mls(ions, rel_perm, v0, s):
    """Instantaneous gate: m = mInf = 0.5*(1 + tanh((V - v0)/s))."""
    return _ml(ions, rel_perm, (TANH, 0.5, 0.5, float(v0), float(s)), _c(1.0), 1.0, False)


def _na_m(v0):          # the Hammil-type activation shared by Nav1p2/1p3/Rat1/Rat3 (vg_na.py:169-186 ...)
    a, b = _linoid(0.182, v0, 9.0), _linoid_neg(0.124, -v0, 9.0)
    return ("R", a, b), ("S", a, b)


_one = _T(_c(1.0))
_m12, _t12 = _na_m(-35.0)
_m13, _t13 = _na_m(-26.0)
_hT_rat = ("S", _linoid(0.024, -50.0, 5.0), _linoid_neg(0.0091, 75.000123, 5.0))
_r2ma, _r2mb = _linoid(0.091, -38.0, 5.0), _linoid_neg(0.062, 38.0, 5.0)
_r2ha, _r2hb = _exp(0.016, -55.0, 15.0), _sig(17.0, -21.0, A=2.07)
_c21a, _c21b = _sig(8.0, -12.5, A=8.5), _sig(-74.0, 14.5, A=35.0)
_c23ma, _c23mb = _sig(-7.0, -8.0, A=2.6), _sig(-26.0, 4.0, A=0.18)
_c23ha, _c23hb = _sig(-32.0, 8.0, A=0.0025), _sig(-42.0, -10.0, A=0.19)
_c22ma, _c22mb = _linoid(0.1, 20.0, 10.0), _exp(0.4, -25.0, 18.0)
_c22ha, _c22hb = _exp(0.01, -50.0, 10.0), _sig(-17.0, -17.0, A=0.1)
_cgma, _cgmb = _linoid(0.055, -27.0, 3.8), _exp(0.94, -75.0, 17.0)
_cgha, _cghb = _exp(0.000457, -13.0, 50.0), _sig(-15.0, -28.0, A=0.0065)

MODELS = {
    # ---- vg_na.py
    "Nav1p2": _hh("Na", 3, 1, _m12, _t12, _T(_sig(-65.0, 6.2)), _hT_rat, shift=-10.0),        # :131-186
    "Nav1p3": _hh("Na", 3, 1, _m13, _t13, _T(_sig(-65.0, 8.1)), _T(_exp(0.265, 0.0, 9.47, c0=0.40))),  # :188-240
    "NavRat2": _hh("Na", 3, 1, ("R", _r2ma, _r2mb), ("S", _r2ma, _r2mb), ("R", _r2ha, _r2hb), ("S", _r2ha, _r2hb)),  # :242-286
    "NavRat1": _hh("Na", 3, 1, _m12, _t12, _T(_sig(-65.0, 6.2)), _hT_rat),                   # :288-328
    "NavRat3": _hh("Na", 3, 1, _m12, _t12, _T(_sig(-65.0, 6.2)), _hT_rat),                   # :330-370
    "NaLeak": _leak("Na"),                                                                    # :372-413
    "Nav1p6": _hh("Na", 1, 0, _T(_sig(-17.0, -1.0 / (0.03937 * 4.2))), _one, _one, _one),     # :415-452
    # ---- vg_k.py
    "Kv1p1": _hh("K", 1, 2, _T(_sig(-30.5, -11.3943)), _T(_sig(-76.56, 26.1479, A=30.0)),
                 _T(_sig(-30.0, 27.3943)), _T(_sig(-160.56, -100.0, A=15000.0))),             # :133-178
    "Kv1p2": _hh("K", 1, 1, _T(_sig(-21.0, -11.3943)), _T(_sig(-67.56, 34.1479, A=150.0)),
                 _T(_sig(-22.0, 11.3943)), _T(_sig(-46.56, -44.1479, A=15000.0))),            # :180-229
    "Kv1p3": _hh("K", 1, 1, _T(_sig(-14.1, -10.3)), _T(_lin(-0.2840, 19.16)),
                 _T(_sig(-33.0, 3.7)), _T(_lin(-13.76, 1162.4))),                             # :231-278
    "Kv1p4": _hh("K", 1, 1, _T(_sig(-21.7, -16.9)), _T(_c(3.0)), _T(_sig(-73.6, 12.8)), _T(_c(119.0))),  # :280-323
    "Kv1p5": _hh("K", 1, 1, _T(_sig(-6.0, -6.4)), _T(_lin(-0.1163, 8.33)),
                 _T(_sig(-25.3, 3.5)), _T(_lin(-15.5, 1620.0))),                              # :325-374
    "Kv1p6": _hh("K", 1, 1, _T(_sig(-20.8, -8.1)), _T(_sig(-46.56, 44.14, A=30.0)),
                 _T(_sig(-22.0, 11.39)), _T(_sig(-46.56, -44.14, A=5000.0))),                 # :376-419
    "Kv2p1": _hh("K", 1, 1, _T(_sig(-9.2, -6.6)), _T(_sig(-46.56, 44.14, A=100.0)),
                 _T(_sig(-19.0, 5.0)), _T(_sig(-46.56, -44.14, A=10000.0))),                  # :421-454
    "Kv2p2": _hh("K", 1, 1, _T(_sig(5.0, -12.0)), _T(_sig(-46.56, -44.14, A=130.0)),
                 _T(_sig(-16.3, 4.8)), _T(_sig(-46.56, -44.14, A=10000.0))),                  # :456-487
    "Kv3p1": _hh("K", 1, 0, _T(_sig(18.7, -9.7)), _T(_sig(-46.56, -44.14, A=20.0)), _one, _one),  # :489-519
    "Kv3p2": _hh("K", 2, 0, _T(_sig(-0.373267, -8.568187)),
                 _T(_sig(19.220623, 4.451533, A=19.106496, c0=3.241643)), _one, _one),        # :521-551
    "Kv3p3": _hh("K", 2, 1, _T(_sig(35.0, -7.3)), _T(_sig(22.414149, 9.704638, A=27.913114, c0=0.676808)),
                 _T(_sig(-28.293856, 29.385636, A=0.75, c0=0.25)),
                 _T(_exp(2776.119438, 0.0, 7.309565, c0=199.786728))),                        # :553-594
    "Kv3p4": _hh("K", 1, 1, _T(_sig(-3.4, -8.4)), _T(_sig(4.44, 38.14, A=10.0)),
                 _T(_sig(-53.32, 7.4)), _T(_sig(-46.56, -44.14, A=20000.0))),                 # :596-639
    "K_Fast": _hh("K", 1, 1, _T(_sig(-47.0, -29.0)), _T(_gauss(0.34, 0.92, 71.0, 59.0)),
                  _T(_sig(-56.0, 10.0)), _T(_gauss(8.0, 49.0, 73.0, 23.0))),                  # :641-682
    "KLeak": _leak("K"),                                                                      # :684-728
    "Kir2p1": _hh("K", 1, 2, _T(_sig(-96.48, 23.26)), _T(_sig(-32.9, 27.93, A=-3.37, c0=3.7)),
                  _T(_sig(-168.28, -44.13)), _T(_sig(-118.29, -27.23, A=306.3, c0=0.85))),    # :730-775
    # ---- vg_ca.py
    "Cav3p3": _hh("Ca", 1, 1, _T(_sig(-45.454426, -5.073015)),
                  _T(_sig(-40.040397, 4.110392, A=54.187616, c0=3.394938)),
                  _T(_sig(-74.031965, 8.416382)), _T(_exp(0.003816, 0.0, 4.781719, c0=109.701136))),  # :132-179
    "Cav3p1": _hh("Ca", 1, 1, _T(_sig(-42.921064, -5.163208)),
                  _T((EXP_CUT, -0.855809, 1.493527, -10e-3, 27.414182)),       # the cut is at -10e-3 *mV*, as written
                  _T(_sig(-72.907420, 4.575763)), _T(_exp(0.002883, 0.0, 5.598574, c0=9.987873))),  # :468-520
    "Cav2p1": _hh("Ca", 1, 0, ("R", _c21a, _c21b), ("S", _c21a, _c21b), _one, _one),          # :181-240
    "Cav1p3": _hh("Ca", 2, 1, _T(_sig(-30.0, -6.0)), _T(_sig(-25.0, 5.0, A=20.0, c0=5.0)),
                  _T(_sig(-80.0, 6.4)), _T(_sig(-40.0, 7.0, A=50.0, c0=20.0))),               # :242-290
    "Cav1p2": _hh("Ca", 2, 1, _T(_sig(-30.0, -6.0)), _T(_sig(-25.0, 5.0, A=20.0, c0=5.0)),
                  _T(_sig(-80.0, 6.4)), _T(_sig(-40.0, 7.0, A=50.0, c0=20.0)), shift=-10.0),  # :292-340
    "Cav2p3": _hh("Ca", 1, 1, ("R", _c23ma, _c23mb), ("S", _c23ma, _c23mb),
                  ("R", _c23ha, _c23hb), ("S", _c23ha, _c23hb)),                              # :342-400
    "Cav2p2": _hh("Ca", 2, 1, ("R", _c22ma, _c22mb), ("S", _c22ma, _c22mb),
                  ("R", _c22ha, _c22hb), ("S", _c22ha, _c22hb)),                              # :402-466
    "Ca_L2": _hh("Ca", 2, 1, _T(_sig(-30.0, -6.0)), _T(_c(10.0)), _T(_sig(-80.0, 6.4)), _T(_c(59.0))),  # :522-573
    "Ca_L3": _hh("Ca", 2, 1, _T(_sig(-30.0, -6.0)), _T(_c(10.0)), _T(_sig(-80.0, 6.4)), _T(_c(59.0)),
                 shift=-15.0),                                                                # :575-635
    "Cav_G": _hh("Ca", 2, 1, ("R", _cgma, _cgmb), ("S", _cgma, _cgmb), ("R", _cgha, _cghb), ("S", _cgha, _cghb)),  # :637-697
    "CaLeak": _leak("Ca"),                                                                    # :699-745
    # ---- vg_cl.py
    "ClLeak": _leak("Cl"),                                                                    # :132-175
    # ---- vg_funny.py: hyperpolarisation-activated Na/K/Ca channels (ions, rel_perm: :73-74)
    "HCN2": _hcn(-99.0, 6.2, 184.0, -10.0),                                                   # :134-191
    "HCN4": _hcn(-100.0, 9.6, 461.0, -10.0),                                                  # :193-251
    "HCN1": _hcn(-94.0, 8.1, 30.0, 0.0),                                                      # :253-289
    "HCNLeak": _multi(_leak("Na"), ["Na", "K", "Ca"], [0.33, 1.0, 0.05], "vg_funny"),         # :291-344
    "HCN2_cAMP": _hcn(-99.0, 6.2, 184.0, -20.0),                                              # :346-404
    "HCN4_cAMP": _hcn(-100.0, 9.6, 461.0, -24.0, init_shift=-20.0),                           # :406-465 (init V-20, run V-24)
    # ---- cation.py: non-selective leaks (ions, rel_perm: :64-65)
    "CatLeak": _multi(_leak("Na"), ["Na", "K", "Ca"], [1.0, 1.0, 0.0], "cation"),             # :113-159
    "CatLeak2": _multi(_leak("Na"), ["Na", "K", "Ca"], [1.0, 1.0, 1.0], "cation"),            # :162-208
    # ---- vg_morrislecar.py (channel class "ML")
    "Kv_ML1": _mlk(12.0, 17.0),                                                               # :119-154
    "Kv2p1_ML": _mlk(14.0, 30.0),                                                             # :156-192
    "Kv1p3_ML": _mlk(-14.0, 20.0),                                                            # :194-230
    "Kv1p5_ML": _mlk(-6.0, 15.0),                                                             # :232-268
    "Kv1p5S_ML": _mlk(-6.0, 15.0, phi=0.00066 * 2),                                           # :270-307
    "Nav_ML": _mls(["Na"], [1], -17.0, 18.0),                                                 # :309-344
    "Cav_L_ML": _mls(["Ca"], [1], -20.0, 24.0),                                               # :346-381
    "Cav_L_ML2": _mls(["Ca"], [1], -20.0, 12.0),                                              # :383-418
    "Cav_N_ML": _mls(["Ca"], [1], -1.0, 18.0),                                                # :420-455
    "Cav_T_ML": _mls(["Ca"], [1], -43.0, 24.0),                                               # :457-492
    "Kir_ML": _ml(["K"], [1], (SECH, 1.0, -135.0, 37.0, 0.0), _c(1.0), 1.0, False),           # :494-529
    "HCN2_ML": _mls(["K", "Na"], [1, 0.2], -99.0, 12.4),                                      # :532-567
    "HCN4_ML": _mls(["K", "Na"], [1, 0.2], -99.0, 19.2),                                      # :569-604
}
# Not tabulated (refused at set-up): the wound channel (channels/wound_channel.py).

CLASS_OF_ION = {"Na": "vg_na", "K": "vg_k", "Ca": "vg_ca", "Cl": "vg_cl"}


def term(t, U):
    ty, p0, p1, p2, p3 = t
    if ty == CONST:
        return np.full_like(U, p0)
    if ty == SIG:
        return p0 + p1 / (1 + np.exp((U - p2) / p3))
    if ty == LIN:
        return p0 * U + p1
    if ty == EXP:
        return p0 + p1 * np.exp((p2 - U) / p3)
    if ty == GAUSS:
        return p0 + p1 * np.exp(-((U + p2) / p3) ** 2)
    if ty == LINOID:
        return p0 * (U - p1) / (1 - np.exp(-(U - p1) / p2))
    if ty == LINOID_NEG:
        W = -U
        return p0 * (W - p1) / (1 - np.exp(-(W - p1) / p2))
    if ty == EXP_CUT:
        return np.where(U >= p2, 1.0, p0 + p1 * np.exp(-U / p3))
    if ty == SECH:
        return p0 / np.cosh((U - p1) / p2)
    if ty == TANH:
        return p0 + p1 * np.tanh((U - p2) / p3)
    raise ValueError(ty)


def quantity(q, U):
    if q[0] == "T":
        return term(q[1], U)
    a, b = term(q[1], U), term(q[2], U)
    return a / (a + b) if q[0] == "R" else 1 / (a + b)


def gates(model, vm):
    """(mInf, mTau, hInf, hTau) of ``model`` at membrane voltages ``vm`` [V] (NumPy; set-up/tests)."""
    M = MODELS[model]
    U = np.asarray(vm, dtype=float) * 1000 + M["shift"]
    return tuple(quantity(q, U) for q in M["q"])


def device_quantities(model):
    """The four quantities as the device kernel takes them: a Morris-Lecar gate becomes a Hodgkin-Huxley gate with the
    time constant mTau/Phi (kinetic) or 0 (instantaneous: (0*m + dt*mInf)/(0 + dt))."""
    M = MODELS[model]
    q = list(M["q"])
    if "ml" in M:
        if M["ml"]["kinetic"]:
            ty, p0, p1, p2, p3 = q[1][1]
            q[1] = _T((ty, p0 / M["ml"]["phi"], p1, p2, p3))
        else:
            q[1] = _T(_c(0.0))
    return q


def advance_gates(model, m, h, vm, dt):
    """One time step of the gates (channelsabc.py:40-71) at membrane voltages ``vm`` [V]; dt = p.dt (NumPy; oracle/tests)."""
    M = MODELS[model]
    mInf, mTau, hInf, hTau = gates(model, vm)
    dtu = dt * M["time_unit"]
    if "ml" in M:
        if M["ml"]["kinetic"]:                     # update_ml, channelsabc.py:70-71
            phi = M["ml"]["phi"]
            return (m + (dtu * phi * mInf / mTau)) / (1 + ((dtu * phi) / mTau)), h
        return mInf * np.ones_like(np.asarray(m, dtype=float)), h      # vg_morrislecar.py:92
    return (mTau * m + dtu * mInf) / (mTau + dtu), (hTau * h + dtu * hInf) / (hTau + dtu)


def initial_state(model, vm):
    """m, h at the first time step (every tabulated model starts at its steady state:
    e.g. vg_na.py:210-228)."""
    M = MODELS[model]
    if "init_shift" in M:
        U = np.asarray(vm, dtype=float) * 1000 + M["init_shift"]
        return quantity(M["q"][0], U), quantity(M["q"][2], U)
    mInf, _, hInf, _ = gates(model, vm)
    return mInf, hInf


def ions_of(model):
    """(ions, rel_perm) a model conducts, in the order the reference applies them."""
    M = MODELS[model]
    return list(M.get("ions", [M["ion"]])), list(M.get("rel_perm", [1.0]))


def make_channel(name, model, max_Dm, targets=None, init_active=True, rel_perm=1.0, m=None, h=None):
    """A channel spec as TissueEngine.set_channels expects it."""
    if model not in MODELS:
        raise KeyError("channel type %r is not tabulated (betse_b200/channels.py)" % model)
    ions, perms = ions_of(model)
    return {"name": name, "model": model, "ion": ions[0], "ions": ions, "maxDm": float(max_Dm),
            "targets": None if targets is None else np.asarray(targets, dtype=np.int64),
            "init_active": bool(init_active), "rel_perm": float(rel_perm) * perms[0],
            "rel_perms": [float(rel_perm) * x for x in perms], "m": m, "h": h}
