// Device-visible plan of the halo exchange between neighbouring strips (include/betse_b200.h,
// "Multi-GPU").  Pointers of a neighbour are addresses inside ITS window, valid in this process
// (CUDA-IPC mapping over NVLink, or the raw pointer when both ranks share a process/device).
#pragma once
#include <stdint.h>

struct XNbr {
    double *cc_mid[2], *vm_cell[2], *flux, *cc_env[2], *v_raw;
    unsigned long long *flags;      // the neighbour's flags, [2 exchange points][2 sides]
    int Cn, En;                     // the neighbour's leading dimensions (cells, env points)
    int side;                       // 0: the neighbour is the strip below, 1: above
    int n_send_cells, recv_cell0;
    int n_send_flux, recv_slot0;
    const int *send_cells, *send_flux;
    const int *send_flux_ell;       // position of the same membranes' fluxes in flux_ell (k_cell), or null
    int cc_rows, cc_src_row0, cc_dst_row0;
    int v_rows, v_src_row0, v_dst_row0;
};

struct XPlan {
    int n_nbr;
    XNbr nb[2];
    unsigned long long *my_flags;   // [2][2] written by the neighbours
    unsigned long long *epoch;      // [2]    exchanges completed on this rank, per exchange point
    unsigned int *done_ctr;         // [2]    CTAs that finished pushing
    unsigned long long timeout_ns;
};
