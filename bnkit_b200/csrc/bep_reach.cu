// bep_reach.cu -- graph assembly of bi-directional edge parsimony on the device.
//
// After the parsimony sweeps (bep.cu) the reference builds, per ancestor, a partial-order graph from the optimal
// edges of all 2 (nPos + 1) instances (POGraph.createFromEdgeMap, src/dat/pog/POGraph.java:544-559), asks whether a
// start-to-end path exists (isContiguous = isPath(-1, N), :314-335) and, if so, removes the nodes that stick out
// (nibble, :566-617): one ascending pass dropping every node without a surviving predecessor that is not a start
// node, one descending pass dropping every node without a surviving successor that is not an end node.  Every
// edge runs from a lower to a higher column, so
//     alive1[c] = start edge into c  OR  an edge p -> c with alive1[p]                       (ascending pass)
//     alive2[c] = alive1[c] AND ( end edge out of c  OR  an edge c -> t with alive2[t] )     (descending pass)
// and the ancestor's presence row IS alive2 (the columns on some start-to-end path); the graph is contiguous iff
// alive2 is non-empty.  r1 did this on the host from 1.6 GB of optimal sets copied back (0.5 s of 0.74 s at C3 size);
// here one THREAD walks one ancestor's columns, 32 ancestors per CTA, with the optimal sets staged through shared
// memory in tiles of 32 columns so that global loads are coalesced rows of B [ancestor][word][instance].
// Only ancestors WITHOUT a path go back to the host (patchAncestorsWithBEP, src/asr/Prediction.java:1531-1590).
#include "bep.cuh"

namespace bnk {

namespace {
constexpr int RT = 32;  // ancestors per CTA = columns per tile

struct ReachArgs {
    const uint32_t *B;     // [n_int][W][n_inst] optimal sets
    const int32_t *dict;   // [sym][n_inst]
    const int32_t *nsym;   // [n_inst]
    int32_t n_int, n_inst, n_pos, W, n_words, mx;   // mx = largest symbol count of an instance
    uint32_t *A1, *Z;      // [n_words][n_int] alive1 / alive2 bit sets (column c = bit c & 31 of word c >> 5)
    uint8_t *contig;       // [n_int]
};

__device__ __forceinline__ bool getbit(const uint32_t *bits, int n_int, int ent, int c)
{
    return (bits[(size_t)(c >> 5) * n_int + ent] >> (c & 31)) & 1u;
}
__device__ __forceinline__ void setbit(uint32_t *bits, int n_int, int ent, int c)
{
    bits[(size_t)(c >> 5) * n_int + ent] |= 1u << (c & 31);
}

// stage the optimal sets of RT ancestors x RT instances (k0 .. k0 + RT - 1) into shared memory [w][anc][RT + 1]:
// 32 x W asynchronous 4-byte copies per lane (LDGSTS), all in flight at once; out-of-range entries are zero-filled
__device__ __forceinline__ void load_tile(const ReachArgs &a, int ent0, int k0, uint32_t *s)
{
    const int lane = threadIdx.x;
    const int k = k0 + lane;
    const bool k_ok = k >= 0 && k < a.n_inst;
    for (int r = 0; r < RT; r++) {
        const int ent = ent0 + r;
        const bool ok = k_ok && ent < a.n_int;
        for (int w = 0; w < a.W; w++) {
            const uint32_t *src = ok ? a.B + ((size_t)ent * a.W + w) * a.n_inst + k : a.B;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(s + (w * RT + r) * (RT + 1) + lane);
            const int bytes = ok ? 4 : 0;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
        }
    }
}
// the symbol counts and dictionary entries of the same RT instances: s_n [RT], s_d [sym][RT + 1]
__device__ __forceinline__ void load_dict_tile(const ReachArgs &a, int k0, int32_t *s_n, int32_t *s_d)
{
    const int lane = threadIdx.x;
    const int k = k0 + lane;
    const bool ok = k >= 0 && k < a.n_inst;
    const int bytes = ok ? 4 : 0;
    {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(s_n + lane);
        const int32_t *src = ok ? a.nsym + k : a.nsym;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    }
    for (int sym = 0; sym < a.mx; sym++) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(s_d + sym * (RT + 1) + lane);
        const int32_t *src = ok ? a.dict + (size_t)sym * a.n_inst + k : a.dict;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    }
}
__device__ __forceinline__ void tiles_ready()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
}

// calls fn(value) for every optimal non-NULL symbol of (this thread's ancestor, tile column j)
struct TileView {   // one direction of the staged tile
    const uint32_t *bits;   // [w][anc][RT + 1]
    const int32_t *nsym;    // [RT]
    const int32_t *dict;    // [sym][RT + 1]
};
template <typename F>
__device__ __forceinline__ void for_each_value(const ReachArgs &a, const TileView &tv, int j, F &&fn)
{
    const int ns = tv.nsym[j];
    if (ns < 1) return;   // nothing to infer: the reference makes no Parsimony object (Prediction.java:1466-1468)
    for (int w = 0; w < a.W; w++) {
        uint32_t bits = tv.bits[(w * RT + threadIdx.x) * (RT + 1) + j];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int sym = w * 32 + b;
            if (sym < ns) fn(tv.dict[sym * (RT + 1) + j]);   // sym == ns is the NULL symbol
        }
    }
}
}  // namespace

// first-pass layout of the instances: forward instance of column c (c = -1 .. nPos - 1) is k = c + 1, backward
// instance of column c (c = 0 .. nPos) is k = nPos + 1 + c
__global__ void __launch_bounds__(RT) bep_reach_kernel(ReachArgs a)
{
    extern __shared__ uint32_t smem[];
    uint32_t *s_f = smem, *s_b = smem + a.W * RT * (RT + 1);
    int32_t *s_nf = reinterpret_cast<int32_t *>(s_b + a.W * RT * (RT + 1)), *s_nb = s_nf + RT;
    int32_t *s_df = s_nb + RT, *s_db = s_df + a.mx * (RT + 1);
    const TileView fw{s_f, s_nf, s_df}, bw{s_b, s_nb, s_db};
    const int ent0 = blockIdx.x * RT, ent = ent0 + threadIdx.x;
    const bool valid = ent < a.n_int;
    const int n_pos = a.n_pos, n_int = a.n_int;
    const int n_tiles = (n_pos + RT - 1) / RT;
    // ---- ascending: alive1.  A1 was zeroed; bit c holds "an alive predecessor / a start edge was seen" until c is visited ----
    load_tile(a, ent0, 0, s_f);   // forward instance of column -1 (k = 0) sits in column 0 of this tile
    load_dict_tile(a, 0, s_nf, s_df);
    tiles_ready();
    if (valid) for_each_value(a, fw, 0, [&](int t) { if (t >= 0 && t < n_pos) setbit(a.A1, n_int, ent, t); });
    __syncwarp();
    for (int T = 0; T < n_tiles; T++) {
        const int c0 = T * RT;
        load_tile(a, ent0, c0 + 1, s_f);           // forward instances of columns c0 .. c0 + 31
        load_tile(a, ent0, n_pos + 1 + c0, s_b);   // backward instances of the same columns
        load_dict_tile(a, c0 + 1, s_nf, s_df);
        load_dict_tile(a, n_pos + 1 + c0, s_nb, s_db);
        tiles_ready();
        if (valid) {
            uint32_t word = a.A1[(size_t)T * n_int + ent];
            for (int j = 0; j < RT && c0 + j < n_pos; j++) {
                bool alive = (word >> j) & 1u;
                if (!alive)
                    for_each_value(a, bw, j, [&](int p) {
                        if (p == -1) alive = true;
                        else if (p >= c0) alive |= ((word >> (p - c0)) & 1u) != 0;
                        else if (p >= 0) alive |= getbit(a.A1, n_int, ent, p);
                    });
                if (alive) {
                    word |= 1u << j;
                    for_each_value(a, fw, j, [&](int t) {
                        if (t >= n_pos) return;
                        if (t < c0 + RT) word |= 1u << (t - c0);
                        else setbit(a.A1, n_int, ent, t);
                    });
                } else {
                    word &= ~(1u << j);
                }
            }
            a.A1[(size_t)T * n_int + ent] = word;
        }
        __syncwarp();
    }
    // ---- descending: alive2.  Z was zeroed; bit c holds "an alive successor / an end edge was seen" until c is visited ----
    load_tile(a, ent0, 2 * n_pos + 1, s_b);   // backward instance of column nPos (the end terminal) in column 0 of the tile
    load_dict_tile(a, 2 * n_pos + 1, s_nb, s_db);
    tiles_ready();
    if (valid) for_each_value(a, bw, 0, [&](int p) { if (p >= 0 && p < n_pos) setbit(a.Z, n_int, ent, p); });
    __syncwarp();
    uint32_t any = 0;
    for (int T = n_tiles - 1; T >= 0; T--) {
        const int c0 = T * RT;
        load_tile(a, ent0, c0 + 1, s_f);
        load_tile(a, ent0, n_pos + 1 + c0, s_b);
        load_dict_tile(a, c0 + 1, s_nf, s_df);
        load_dict_tile(a, n_pos + 1 + c0, s_nb, s_db);
        tiles_ready();
        if (valid) {
            uint32_t word = a.Z[(size_t)T * n_int + ent];
            const uint32_t alive1 = a.A1[(size_t)T * n_int + ent];
            const int jhi = (n_pos - c0 < RT ? n_pos - c0 : RT) - 1;
            for (int j = jhi; j >= 0; j--) {
                bool ok = false;
                if ((alive1 >> j) & 1u) {
                    ok = (word >> j) & 1u;
                    if (!ok)
                        for_each_value(a, fw, j, [&](int t) {
                            if (t == n_pos) ok = true;
                            else if (t < c0 + RT) ok |= ((word >> (t - c0)) & 1u) != 0;
                            else if (t < n_pos) ok |= getbit(a.Z, n_int, ent, t);
                        });
                }
                if (ok) {
                    word |= 1u << j;
                    for_each_value(a, bw, j, [&](int p) {
                        if (p < 0) return;
                        if (p >= c0) word |= 1u << (p - c0);
                        else setbit(a.Z, n_int, ent, p);
                    });
                } else {
                    word &= ~(1u << j);
                }
            }
            a.Z[(size_t)T * n_int + ent] = word;
            any |= word;
        }
        __syncwarp();
    }
    if (valid) a.contig[ent] = any != 0;
}

// present[node][col]: extants from their residues, ancestors from alive2 (rows of ancestors without a path are
// overwritten by the host after its patch rounds)
__global__ void bep_rows_kernel(const uint8_t *__restrict__ obs, const int32_t *__restrict__ ent_of_node, const uint32_t *__restrict__ Z,
                                int n_nodes, int n_int, int n_pos, uint8_t *__restrict__ present)
{
    const int node = blockIdx.x;
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= n_pos) return;
    const int ent = ent_of_node[node];
    uint8_t v;
    if (ent < 0) v = obs[(size_t)node * n_pos + c] != BNK_NONE;
    else v = (Z[(size_t)(c >> 5) * n_int + ent] >> (c & 31)) & 1u;
    present[(size_t)node * n_pos + c] = v;
}

// rows[i] of B gathered into a compact buffer [n_rows][row_words]
__global__ void bep_gather_kernel(const uint32_t *__restrict__ B, const int32_t *__restrict__ rows, int n_rows, size_t row_words,
                                  uint32_t *__restrict__ out)
{
    const int r = blockIdx.x;
    const uint32_t *src = B + (size_t)rows[r] * row_words;
    uint32_t *dst = out + (size_t)r * row_words;
    for (size_t i = (size_t)blockIdx.y * blockDim.x + threadIdx.x; i < row_words; i += (size_t)gridDim.y * blockDim.x) dst[i] = src[i];
}

cudaError_t bep_launch_reach(const uint32_t *B, const int32_t *dict, const int32_t *nsym, int mx, int n_int, int n_inst, int n_pos, int W,
                             uint32_t *A1, uint32_t *Z, uint8_t *contig, cudaStream_t st)
{
    ReachArgs a;
    a.B = B; a.dict = dict; a.nsym = nsym; a.n_int = n_int; a.n_inst = n_inst; a.n_pos = n_pos; a.W = W; a.mx = mx < 1 ? 1 : mx;
    a.n_words = (n_pos + 31) / 32; a.A1 = A1; a.Z = Z; a.contig = contig;
    cudaError_t e = cudaMemsetAsync(A1, 0, sizeof(uint32_t) * (size_t)a.n_words * n_int, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(Z, 0, sizeof(uint32_t) * (size_t)a.n_words * n_int, st);
    if (e != cudaSuccess) return e;
    const size_t smem = sizeof(uint32_t) * (2 * (size_t)W * RT * (RT + 1) + 2 * RT + 2 * (size_t)a.mx * (RT + 1));
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(bep_reach_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    bep_reach_kernel<<<(n_int + RT - 1) / RT, RT, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t bep_launch_rows(const uint8_t *obs, const int32_t *ent_of_node, const uint32_t *Z, int n_nodes, int n_int, int n_pos,
                            uint8_t *present, cudaStream_t st)
{
    bep_rows_kernel<<<dim3(n_nodes, (n_pos + 255) / 256), 256, 0, st>>>(obs, ent_of_node, Z, n_nodes, n_int, n_pos, present);
    return cudaGetLastError();
}

cudaError_t bep_launch_gather(const uint32_t *B, const int32_t *rows, int n_rows, size_t row_words, uint32_t *out, cudaStream_t st)
{
    if (n_rows == 0) return cudaSuccess;
    bep_gather_kernel<<<dim3(n_rows, 32), 256, 0, st>>>(B, rows, n_rows, row_words, out);
    return cudaGetLastError();
}

}  // namespace bnk
