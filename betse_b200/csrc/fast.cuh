// Device-side state of the fast (equivalent-circuit) solver (csrc/fast.cu).
#pragma once

struct KFast {
    const double *G_Leak, *E_Leak, *G_gj, *sigma_cell;     // [C]
    const double *extra_J;                                  // [M] sim.extra_J_mem, or null
    double *vm_ave[2];                                      // [C] double buffered
    double *vgj, *Jn, *Emx, *Emy;                           // [M]
    double *J_cell_x, *J_cell_y, *E_cell_x, *E_cell_y;      // [C]
    double sm;                                              // 0.1*sim.sigma_cell.mean()
    double dt_cm;                                           // p.dt*(1/p.cm)
};
