// Device-side description of one network handler (general network / gene regulatory network):
// substances, compiled rate-law programs (betse_b200/ratelaw.py) and their tables.
#pragma once
#include "kparams.cuh"

#define NET_MAX_RATES 48    // growth/decay rates + cell-zone reactions of one handler
#define NET_STACK 16        // ratelaw.MAX_STACK

// opcodes of the postfix programs (ratelaw.py)
enum { RL_PUSHC = 0, RL_PUSHS, RL_PUSHA, RL_PUSHI, RL_PUSHM, RL_PUSHV, RL_PUSHE, RL_PUSHJ, RL_ADD, RL_SUB, RL_MUL, RL_DIV, RL_POW, RL_NEG, RL_EXP };
#define RL_LAST_PUSH RL_PUSHJ

struct KNet {
    int K;                      // substances
    int E;                      // env points (ny*nx)
    const int* cell2ecm;        // [C] cells.map_cell2ecm (cell-zone rate laws that read concentrations outside the cells)
    int n_rates;                // K growth/decay rates + R reactions (columns of reaction_matrix)
    double* c;                  // [K][C] concentrations in the cells
    double* rates;              // [n_rates][C] rates of the last step (reaction_rates / download)
    double* gj_delta;           // [C] scratch of the gap-junction transport
    const int* code;            // (op, arg) pairs
    const int* ptr;             // [n_programs + 1] program ranges (in pairs)
    const double* consts;
    const double* cell_arrays;  // [n][C] per-cell constants (growth_mod_function_cells ...)
    const double* mem_arrays;   // [n][M] per-membrane constants
    const unsigned char* gmask; // [K][C] growth_targets_cell (null: every cell)
    const double* stoich;       // [K][n_rates] substance rows of reaction_matrix
    const double* Dgj;          // [K] gap-junction diffusion constant; < 0: ignoreGJ
    const double* z;            // [K]
    const double* tdf;          // [K] modify_time_factor
    // membrane / extracellular legs (null when no substance of the handler has them)
    double* c_env;              // [K][E] env concentrations
    double* env_tmp;            // [E] scratch of the env transport
    double* mem_delta;          // [K][C] sum_mems(f_mem*mem_sa)/cell_vol of this step's membrane flux
    const double* Dm;           // [K]
    const double* c_bound;      // [K]
    const double* D_env;        // [K][E]
    const int* env_rx_prog;     // [n_env_rx] extracellular reactions: program indices
    int n_env_rx;
    const double* stoich_env;   // [K][n_env_rx] substance rows of reaction_matrix_env
    const unsigned char* env_on_d; // [K] the substance exists outside the cells
    // Membrane values (mem_concs[X] = Molecule.cc_at_mem / sim.cc_at_mem[ion]) of everything a transporter moves in the
    // cells: the reference updates the cell value AND nudges the membrane value (networks.py:3016-3022), two separate
    // arrays until the next update_intra / update_Co — so later membrane-zone rate laws of the step read these rows
    // (gathered from the cells when the handler's block starts, re-gathered when a channel's update_Co renews an ion)
    double* tw;                 // [n_rows][M]
    double* tr_flux;            // [M] flux of the transporter in flight
    const double* sa_over_vol;  // [M] mem_sa/mem_vol
    signed char tw_s[NET_MAX_RATES];   // row of substance k (< 0: none)
    signed char tw_i[8];               // row of ion i
    // 'update intracellular' (diagnostic case): membrane values of their own
    double* cmem;               // [K][M]
    const unsigned char* intra; // [K]
    const double* Do;           // [K]
    const double* R_rads;       // [M]
    const double* mu_mem;       // [K] Molecule.Mu_mem
    double* gjf;                // [M] gap-junction flux of the substance in flight (membrane values move after all are read)
    double* clamp;              // [K] cell clamp in force this step (NaN: none), Molecule.cell_clamp_method
    const unsigned char* pumped; // [K] 1: the substance has its own pump -> its membrane leg follows the pump (launch_net)
    double* c_save;             // [n_pumps][C] a pumped substance's concentration before growth/decay (cc_at_mem of its membrane leg)
    // p.substances_affect_charge (networks.py:2942-2977)
    int affect;
    const double* scale;        // [K] scale_factor
    double* fmem_tmp;           // [M] membrane flux of the substance in flight (joined with its gap-junction flux)
    double* rho_cells;          // [C] extra_rho_cells of this handler
    double* rho_env;            // [E] extra_rho_env
};

// One rate law at cell c (membrane m < 0: cell zone).
__device__ __forceinline__ double rl_eval(const KNet& N, const int prog, const int c, const int m, const KArrays& A,
                                          const int C, const int M, const int cur, const double vm)
{
    double st[NET_STACK];
    int sp = 0;
    const int p0 = __ldg(N.ptr + prog), p1 = __ldg(N.ptr + prog + 1);
    for (int pc = p0; pc < p1; ++pc) {
        const int2 ins = __ldg(reinterpret_cast<const int2*>(N.code) + pc);
        const int op = ins.x, arg = ins.y;
        if (op <= RL_LAST_PUSH) {
            double v;
            switch (op) {
                // outside the cells: substance / ion at the membrane's env square (membrane zone) or at the env square of
                // the cell centre (cell zone, cells.map_cell2ecm); the ions' env
                // field is the one this step's transport left (cc_env[cur ^ 1]), what sim.cc_env holds during the network block
                // (m < -1: an extracellular-zone program — tight-junction modulators — evaluated at env square -2 - m)
                case RL_PUSHE: v = N.c_env[(size_t)arg * N.E + (m >= 0 ? __ldg(A.map_mem2ecm + m) : (m < -1 ? -2 - m : __ldg(N.cell2ecm + c)))]; break;
                case RL_PUSHJ: v = A.cc_env[cur ^ 1][(size_t)arg * N.E + (m >= 0 ? __ldg(A.map_mem2ecm + m) : (m < -1 ? -2 - m : __ldg(N.cell2ecm + c)))]; break;
                case RL_PUSHC: v = __ldg(N.consts + arg); break;
                case RL_PUSHS: v = (m >= 0 && N.tw && N.tw_s[arg] >= 0) ? N.tw[(size_t)N.tw_s[arg] * M + m] : N.c[(size_t)arg * C + c];
                               break;
                case RL_PUSHA: v = (m < 0) ? __ldg(N.cell_arrays + (size_t)arg * C + c) : __ldg(N.mem_arrays + (size_t)arg * M + m); break;
                case RL_PUSHI: v = (m >= 0 && N.tw && N.tw_i[arg] >= 0) ? N.tw[(size_t)N.tw_i[arg] * M + m] : A.cc_cells[(size_t)arg * C + c];
                               break;
                case RL_PUSHM: v = A.cc_mid[cur][(size_t)arg * C + c]; break;
                default: v = vm; break;
            }
            st[sp++] = v;
        } else if (op == RL_NEG) st[sp - 1] = -st[sp - 1];
        else if (op == RL_EXP) st[sp - 1] = exp(st[sp - 1]);
        else {
            const double b = st[--sp], a = st[sp - 1];
            double r;
            switch (op) {
                case RL_ADD: r = a + b; break;
                case RL_SUB: r = a - b; break;
                case RL_MUL: r = a * b; break;
                case RL_DIV: r = a / b; break;
                default:     // NumPy's scalar-exponent fast paths (x**1, x**2, x**0.5, x**-1, x**0), else pow
                    r = (b == 1.0) ? a : (b == 2.0) ? a * a : (b == 0.5) ? sqrt(a) : (b == -1.0) ? 1.0 / a : (b == 0.0) ? 1.0 : pow(a, b);
                    break;
            }
            st[sp - 1] = r;
        }
    }
    return st[0];
}
