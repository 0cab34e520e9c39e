// march_emu.cpp -- TEST INFRASTRUCTURE: the march kernel of mf-lbm_b200/csrc/march.cuh compiled by g++ as a sequential
// emulation (MARCH_EMU), so that its index logic and arithmetic can be checked against the oracle where no GPU exists
// (tests/test_march_emu.py).  Builds the same device-layout inputs mflbm_upload / kernels_march.cu build (padded grid,
// active-index map, cell codes) from the caller's arrays and runs every block of the launch one after the other.
// g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC -DMARCH_EMU -D__host__= -D__device__= -D__global__= -I/usr/local/cuda/include
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../mf-lbm_b200/csrc/march.cuh"

using namespace mflbm;

extern "C" int march_emu_run(int nx, int ny, int nz, const int8_t *walls /* (-1:n+2)^3, i fastest */,
                             const double *phi /* (-3:n+4)^3 */, const mflbm_solid_node *solid, int ns,
                             const mflbm_fluid_node *fluid, int nf, double gamma, int lz, double *G /* [4][nA] */, int nA_expect,
                             int *wstamp_out /* [ceil(nA/32)] or null */, int *flags /* [4]: code errors, blocks, ticks, nA */) {
    Dev P;
    memset(&P, 0, sizeof(P));
    Grid &g = P.g;
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.sx = (nx + 8 + 15) / 16 * 16;
    g.sxy = g.sx * (ny + 8);
    g.base = 16;
    g.ntot = 16 + g.sxy * (nz + 8) + 16;
    g.set_magic();
    const size_t ntot = (size_t)g.ntot;
    std::vector<int8_t> W(ntot, 0);
    std::vector<double> PH(ntot, 0.0);
    std::vector<int> SM(ntot, -1);
    std::vector<unsigned> CODE(ntot, MCODE_NONE);
    for (int k = -1; k <= nz + 2; k++)
        for (int j = -1; j <= ny + 2; j++)
            for (int i = -1; i <= nx + 2; i++)
                W[g.cell(i, j, k)] = walls[(size_t)(i + 1) + (size_t)(nx + 4) * ((size_t)(j + 1) + (size_t)(ny + 4) * (size_t)(k + 1))];
    for (int k = -3; k <= nz + 4; k++)
        for (int j = -3; j <= ny + 4; j++)
            for (int i = -3; i <= nx + 4; i++)
                PH[g.cell(i, j, k)] = phi[(size_t)(i + 3) + (size_t)(nx + 8) * ((size_t)(j + 3) + (size_t)(ny + 8) * (size_t)(k + 3))];
    // row paddings hold garbage on the device: make sure nothing depends on them
    for (int k = -3; k <= nz + 4; k++)
        for (int j = -3; j <= ny + 4; j++)
            for (int o = nx + 8; o < g.sx; o++) PH[g.base - 4 + o + g.sx * (j + 3) + (size_t)g.sxy * (k + 3)] = 1e300;
    int nA = 0;
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++)
                if (W[g.cell(i, j, k)] == 0) SM[g.cell(i, j, k)] = nA++;
    int nS = nA;  // storage-only nodes of the ghost planes carry indices >= nA
    for (int k = 0; k <= nz + 1; k += nz + 1)
        for (int j = 0; j <= ny + 1; j++)
            for (int i = 0; i <= nx + 1; i++)
                if (W[g.cell(i, j, k)] == 0) SM[g.cell(i, j, k)] = nS++;
    flags[3] = nA;
    if (nA != nA_expect) return -1;
    // cell codes, like k_mcode_base / k_mcode_solid / k_mcode_fluid
    int err = 0;
    for (size_t c = 0; c < ntot; c++) CODE[c] = m_code_base(g, W.data(), (long long)c);
    for (int n = 0; n < ns; n++) {
        const mflbm_solid_node &s = solid[n];
        unsigned m = 0;
        for (int t = 0; t < s.i_fluid_num; t++) m |= 1u << s.neighbor_list[t];
        if (s.ix >= 0 && s.ix <= nx + 1 && s.iy >= 0 && s.iy <= ny + 1 && s.iz >= 0 && s.iz <= nz + 1) m |= 0x80000000u;
        const int c = g.cell(s.ix, s.iy, s.iz);
        if (CODE[c] != MCODE_NONE) err |= 1;
        CODE[c] = m_code_solid(m);
        const int na = __builtin_popcount(m & 0x7eu), nb = __builtin_popcount(m & 0x7ff80u);
        const double law = m_law_from_counts(na, nb);
        if (memcmp(&law, &s.la_weight, 8) != 0) err |= 2;
    }
    std::vector<double> NW((size_t)5 * (nf > 0 ? nf : 1));
    for (int n = 0; n < nf; n++) {
        const mflbm_fluid_node &s = fluid[n];
        const int c = g.cell(s.ix, s.iy, s.iz);
        if (CODE[c] != MCODE_FLUID) err |= 4;
        CODE[c] = m_code_fluid((unsigned)n);
        NW[n] = s.nwx; NW[(size_t)nf + n] = s.nwy; NW[2 * (size_t)nf + n] = s.nwz;
        NW[3 * (size_t)nf + n] = cos(s.theta); NW[4 * (size_t)nf + n] = sin(s.theta);
    }
    flags[0] = err;
    std::vector<int> WS((size_t)(nA + 31) / 32 + 1, 0);
    P.phi = PH.data();
    P.walls = W.data();
    P.smap = SM.data();
    P.mcode = CODE.data();
    P.fluid_nw = NW.data();
    P.num_fluid = nf;
    P.nA = nA;
    P.gamma = gamma;
    P.wstamp = WS.data();
    for (int q = 0; q < 4; q++) P.G[q] = G + (size_t)q * nA;
    const int ncx = (nx + MARCH_TX - 1) / MARCH_TX, ncy = (ny + MARCH_TY - 1) / MARCH_TY, nch = (nz + lz - 1) / lz;
    long long ticks = 0;
#pragma omp parallel
    {
        MarchSmem *S = (MarchSmem *)malloc(sizeof(MarchSmem));
        memset(S, 0xA5, sizeof(MarchSmem));  // shared memory starts as garbage
#pragma omp for collapse(2) schedule(dynamic) reduction(+ : ticks)
        for (int ch = 0; ch < nch; ch++)
            for (int b = 0; b < ncx * ncy; b++) {
                const int kA = 1 + ch * lz, kB = (kA + lz - 1 < nz) ? kA + lz - 1 : nz;
                march_block(P, *S, b % ncx, b / ncx, kA, kB, 7);
                ticks += kB - kA + 10;
            }
        free(S);
    }
    flags[1] = ncx * ncy * nch;
    flags[2] = (int)ticks;
    if (wstamp_out) memcpy(wstamp_out, WS.data(), ((size_t)(nA + 31) / 32) * sizeof(int));
    return 0;
}

extern "C" int march_emu_smem_bytes(void) { return (int)sizeof(MarchSmem); }
