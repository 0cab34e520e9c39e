// kernels_march.cu -- the march kernel (march.cuh: K3..K7 of the colour-gradient chain fused on chip, sparse multiphase
// layout): device wrapper, the per-cell codes it reads, and its launch.  DESIGN.md "March kernel".
#include <algorithm>

#include "march.cuh"

namespace mflbm {

// ---- cell codes (once per upload of walls / node lists) ----
__global__ void __launch_bounds__(256) k_mcode_base(const Dev P) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < P.g.ntot) P.mcode[c] = m_code_base(P.g, P.walls, c);
}

// listed solid boundary nodes: mask + box flag; the caller's la_weight must be the reference's sum over the listed
// neighbours (the kernel recomputes it from the mask), and a node may be listed only once and only on a solid / outside cell
__global__ void __launch_bounds__(256) k_mcode_solid(const Dev P, int *err) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P.num_solid) return;
    const unsigned m = P.solid_mask[n];
    const unsigned old = atomicExch(&P.mcode[P.solid_cell[n]], m_code_solid(m));
    if (old != MCODE_NONE) atomicOr(err, 1);
    const double law = m_law_from_counts(__popc(m & 0x7eu), __popc(m & 0x7ff80u));
    if (__double_as_longlong(law) != __double_as_longlong(P.solid_law[n])) atomicOr(err, 2);
}

__global__ void __launch_bounds__(256) k_mcode_fluid(const Dev P, int *err) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P.num_fluid) return;
    const unsigned old = atomicExch(&P.mcode[P.fluid_cell[n]], m_code_fluid((unsigned)n));
    if (old != MCODE_FLUID) atomicOr(err, 4);
}

// ---- the kernel: persistent blocks draw (column, chunk) work items from a ticket counter, in index order, so that the
// blocks in flight work on neighbouring columns of the same planes and share their halos through L2 ----
// mode 0: every item; mode 1: every item, only when the tile update selected the flat sweep (tcount[2]); mode 2: only the
// items listed by the tile update (mlist, tcount[10] of them; those holding an active tile), when it selected the tile-driven pass
__global__ void __launch_bounds__(MARCH_NT, 1) k_march(const Dev P, const int lz, const int mode, const int stamp) {
    extern __shared__ __align__(16) unsigned char march_smem_raw[];
    MarchSmem &S = *reinterpret_cast<MarchSmem *>(march_smem_raw);
    __shared__ int s_item;
    if (mode == 1 && !P.tcount[2]) return;
    if (mode == 2 && !P.tcount[3]) return;
    const int ncol = P.mcols_x * P.mcols_y;
    const int nchunk = (P.g.nz + lz - 1) / lz;
    const int count = mode == 2 ? P.tcount[10] : ncol * nchunk;
    for (;;) {
        __syncthreads();  // the previous item's last phase is over: shared memory (and s_item) may be reused
        if (threadIdx.x == 0) s_item = atomicAdd(&P.tcount[11], 1);
        __syncthreads();
        const int w = s_item;
        if (w >= count) return;
        const int item = mode == 2 ? P.mlist[w] : w;
        const int col = item % ncol, ch = item / ncol;
        int ch1 = ch;
        if (mode == 2) {
            // listed items on top of each other in one column are marched through in one go by the lowest of them: the six
            // ramp-up and four run-out planes are paid once per run instead of once per item
            if (ch > 0 && P.mflag[item - ncol] == stamp) continue;
            while (ch1 + 1 < nchunk && P.mflag[item + (ch1 + 1 - ch) * ncol] == stamp) ch1++;
        }
        const int kA = 1 + ch * lz;
        const int kB = min(P.g.nz, (ch1 + 1) * lz);
        march_block(P, S, col % P.mcols_x, col / P.mcols_x, kA, kB, mode == 2 ? stamp : 0);
    }
}

// (re)builds the cell codes; returns 0 when the march kernel may run, 1 when the node lists are not what it assumes
// (the caller then keeps the list kernels), -1 on a CUDA error
int march_prepare(mflbm_ctx *c, cudaStream_t st) {
    const Dev &P = c->d;
    if (!c->march_on || c->march_ready) return 0;
    int *err = P.tcount + 7;  // scratch word
    cudaMemsetAsync(err, 0, sizeof(int), st);
    k_mcode_base<<<(P.g.ntot + 255) / 256, 256, 0, st>>>(P);
    if (P.num_solid > 0) k_mcode_solid<<<(P.num_solid + 255) / 256, 256, 0, st>>>(P, err);
    if (P.num_fluid > 0) k_mcode_fluid<<<(P.num_fluid + 255) / 256, 256, 0, st>>>(P, err);
    c->launches += 3;
    int h = -1;
    if (cudaMemcpyAsync(&h, err, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return -1;
    cudaMemsetAsync(err, 0, sizeof(int), st);
    if (h != 0) {
        c->march_on = false;  // duplicate / misplaced list entries or a foreign la_weight: the list kernels take them as they are
        c->march_reject = h;
        return 1;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_march, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmem)) != cudaSuccess) return -1;
        attr_set = true;
    }
    c->march_ready = true;
    return 0;
}

static int march_grid() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms < 1) sms = 148;
    }
    return sms;  // one 512-thread block with ~190 KB of shared memory per SM
}

// mode as in k_march; lz_flat planes per work item for the sweeps over everything, P.march_lz for the listed items
void launch_march(mflbm_ctx *c, cudaStream_t st, int mode, int stamp) {
    const Dev &P = c->d;
    const int lz = mode == 2 ? P.march_lz : c->march_lz_flat;
    const int items = P.mcols_x * P.mcols_y * ((P.g.nz + lz - 1) / lz);
    k_march<<<std::min(march_grid(), items), MARCH_NT, sizeof(MarchSmem), st>>>(P, lz, mode, stamp);
    c->launches++;
}

}  // namespace mflbm
