// Device-side description of one voltage-gated channel (csrc/channels.cu, betse_b200/channels.py).
#pragma once

#define KCH_PACK 4           // channels of one pass (channels.cu:k_chan_cell)

struct KTerm { int type; double p[4]; };

struct KChan {
    int ion, mpow, hpow;
    int kind[4];                 // mInf, mTau, hInf, hTau: 0 = term a, 1 = a/(a+b), 2 = 1/(a+b)
    KTerm a[4], b[4];
    double dt_tu;                // p.dt * time_unit
    double maxDm, rel_perm, shift;
    const unsigned char* mask;   // [M] targets (null = every membrane)
    double *m, *h, *P, *flux;    // [M] gate states, open probability, last flux
    double *D;                   // [M] DChan = P*rel_perm*maxDm*moddy (networks.py:3164)
    double *mc, *hc, *Pc, *Dc;   // [C] the same per cell while the channels run on the per-cell path (k_chan_cell)
    double *fell;                // [rows][32] last flux in the order of the cell pack
    int frozen;                  // this entry is a further conducted ion of the PREVIOUS entry's channel (multi-ion families:
                                 // vg_funny, cation): same m/h/P arrays, gates already advanced (networks.py:3156-3158)
    int handler;                 // network handler the channel belongs to (0 general network, 1 gene network)
    int mod_prog;                // program of alpha_eval_string in that handler's table (< 0: moddy == 1)
};
