// Buffers of the Helmholtz-Hodge diagnostics (csrc/hh.cu).
#pragma once

struct HHBuf {
    double *Sy, *Sx;          // [my][my], [mx][mx] sine matrices sin(pi (j+1)(k+1)/(n+1))
    double *ly, *lx;          // eigenvalues 2cos(pi k/(n+1)) - 2
    double *Jx, *Jy;          // [E] total current density of the ions
    double *bA, *bB;          // [E] right-hand sides of the two Poisson problems
    double *uA, *uB;          // [E] potentials AA, BB
    double *R, *T1;           // [2][my][mx] work (two Poisson solves at once)
    double *PLy, *PRx;        // packed even / odd sine matrices (hh.cu: dst_left / dst_right): PLy [hy][my] = [Se | So], PRx [mx][hx]
    double *EO;               // [2 solves][even, odd][max(hy*mx, my*hx)] work of the folded transforms
    double *J_env_x, *J_env_y, *B_field, *Jtx, *Jty;   // outputs [E]
    double mu, bound[4];      // p.mu; sim.bound_V T, B, L, R
};
