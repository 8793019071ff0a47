// The FAST (equivalent-circuit) solver: the body of Simulator._run_fast_sim_core_loop (betse/science/sim.py:1547-1592)
// without networks — a leak circuit per cell (G_Leak, E_Leak = vm_GHK from Simulator.fast_sim_init, sim.py:1393-1452)
// coupled through gap junctions of conductance G_gj, with the voltage-sensitive gating of channels/gap_junction.py:53-77.
//
//   k_fast       one thread per cell: the transjunctional voltages of the cell's membranes from the OLD cell potentials
//                (Jacobi, like the reference's whole-array update), their gating, the two segmented sums
//                (np.dot(cells.M_sum_mems, .)), the new potential; NaN check (stb.check_v)
//   k_fast_diag  sampled steps: Jn, the cell-centre currents and fields, and the per-cell divergence-free membrane field
//                (cells.single_cell_div_free, cells.py:2614-2623)
//
// HBM-bound by construction: per cell 2 x 8 B potential, 3 x 8 B constants; per membrane 4 B index, 3 x 8 B (gjopen r/w,
// vgj) (+ 16 B when a network current is present).
#include <stdint.h>
#include "kparams.cuh"
#include "kmath.cuh"
#include "fast.cuh"


__global__ void __launch_bounds__(256)
k_fast(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KFast Fz, const int cur)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cells_owned) return;
    const double* __restrict__ v = Fz.vm_ave[cur];
    const int m0 = ldgi(A.cell_mem_ptr + c), m1 = ldgi(A.cell_mem_ptr + c + 1);
    const double vc = v[c];
    double S = 0.0, Jm = 0.0;
    for (int m = m0; m < m1; ++m) {
        const int cn = ldgi(A.nn_cell_flag + m) & 0x7fffffff;         // cells.cell_nn_i[m, 1] (the own cell at the boundary)
        const double vgj = v[cn] - vc;                                 // sim.py:1547
        Fz.vgj[m] = vgj;
        const double gjb = A.gj_block ? ldg(A.gj_block + m) : P.gj_block;
        double g;
        if (P.v_sensitive_gj) {
            // Gap_Junction.run (gap_junction.py:53-77): implicit Euler sub-step, written as g' = g*c1 + c2
            double c1, c2;
            gj_gate_map(vgj, P, gjb, c1, c2);
            g = fma(A.gjopen[m], c1, c2);
        } else g = gjb * ldg(A.gj_w + m);                              // sim.py:1555
        A.gjopen[m] = g;
        S += vgj;
        if (Fz.extra_J) Jm += ldg(Fz.extra_J + m) * ldg(A.mem_sa + m);
    }
    const double Jgj = ldg(Fz.G_gj + c) * S;                           // sim.py:1557
    const double Jmem = Jm / ldg(A.cell_sa + c);                       // sim.py:1559
    const double vn = vc + Fz.dt_cm * (Jgj - Jmem - ldg(Fz.G_Leak + c) * (vc - ldg(Fz.E_Leak + c)));   // sim.py:1561
    if (vn != vn) atomicOr(A.status, ST_NAN_VM);                       // stb.check_v(self.vm_ave), sim.py:1592
    Fz.vm_ave[cur ^ 1][c] = vn;
}

__global__ void __launch_bounds__(256)
k_fast_diag(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KFast Fz)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cells_owned) return;
    const int m0 = ldgi(A.cell_mem_ptr + c), m1 = ldgi(A.cell_mem_ptr + c + 1);
    const double Gg = ldg(Fz.G_gj + c);
    double div = 0.0, jx = 0.0, jy = 0.0;
    for (int m = m0; m < m1; ++m) {
        const double Jn = -Fz.vgj[m] * Gg + (Fz.extra_J ? ldg(Fz.extra_J + m) : 0.0);   // sim.py:1566-1568
        Fz.Jn[m] = Jn;
        const double nx = ldg(A.mem_nx + m), ny = ldg(A.mem_ny + m), sa = ldg(A.mem_sa + m);
        const double Jmx = Jn * nx, Jmy = Jn * ny;
        const double ux = Jmx / Fz.sm, uy = Jmy / Fz.sm;               // sim.py:1573
        div += (ux * nx + uy * ny) * sa;                               // cells.div, cells.py:2391-2396
        jx += Jmx * sa; jy += Jmy * sa;                                // sim.py:1579-1580
    }
    const double vol = ldg(A.cell_vol + c), csa = ldg(A.cell_sa + c);
    const double divU = div / vol;
    const double nm = (double)(m1 - m0);
    for (int m = m0; m < m1; ++m) {
        const double nx = ldg(A.mem_nx + m), ny = ldg(A.mem_ny + m), sa = ldg(A.mem_sa + m);
        const double Jn = Fz.Jn[m];
        // M_sum_mems_inv (a pseudo-inverse, cells.py:1308) spreads a cell value over its membranes, divided by their number
        const double Pi = (divU / nm) * (vol / sa);                    // cells.py:2617-2618
        Fz.Emx[m] = (Jn * nx) / Fz.sm - Pi * nx;
        Fz.Emy[m] = (Jn * ny) / Fz.sm - Pi * ny;
    }
    const double Jcx = jx / csa, Jcy = jy / csa;
    const double sg = 0.1 * ldg(Fz.sigma_cell + c);
    Fz.J_cell_x[c] = Jcx; Fz.J_cell_y[c] = Jcy;
    Fz.E_cell_x[c] = Jcx / sg; Fz.E_cell_y[c] = Jcy / sg;              // sim.py:1583-1584
}

void launch_fast(const KParams& P, const KArrays& A, const KFast& Fz, int cur, cudaStream_t st)
{
    k_fast<<<(P.n_cells_owned + 255) / 256, 256, 0, st>>>(P, A, Fz, cur);
}

void launch_fast_diag(const KParams& P, const KArrays& A, const KFast& Fz, cudaStream_t st)
{
    k_fast_diag<<<(P.n_cells_owned + 255) / 256, 256, 0, st>>>(P, A, Fz);
}
