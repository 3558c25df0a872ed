// tu_doudizhu.cu -- kernel instantiations for DouDizhu (warp per env) + the constant action table.
#include <cstring>
#include "game_doudizhu.cuh"
namespace rlc {

// blob written by rlcard_b200/doudizhu_table.py build_blob(): header + 16-byte aligned arrays
struct DdzBlobHeader {
    char magic[4]; uint32_t n_actions, n_words_padded, n_types;
    uint64_t off_rows, off_need, off_type, off_weight, off_tw, total;
};
static DdzTables g_tab[64];
static void *g_dev_blob[64];
static uint16_t *g_therm[64];                 // per-action 54-byte obs blocks, built on the device from the rows at upload
static_assert(sizeof(DdzTables) == 6 * sizeof(void *), "DdzTables is copied into KParams::tab[6]");

// envs/doudizhu.py:153-167 _cards2array for every action id: byte 4 r + k = (count of rank r > k), bytes 52 / 53 = the jokers
__global__ void k_ddz_build_therm(const uint64_t *rows, int n_actions, uint16_t *therm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_actions * 27) return;
    const int id = i / 27, u = i - 27 * id;
    const uint64_t c = rows[id];
    uint32_t v = 0;
    for (int h = 0; h < 2; h++) {
        const int j = 2 * u + h;
        const uint32_t on = j < 52 ? (uint32_t)(((c >> (4 * (j >> 2))) & 15ull) > (uint64_t)(j & 3)) : (uint32_t)(((c >> (4 * (j - 39))) & 15ull) != 0ull);
        v |= on << (8 * h);
    }
    therm[i] = (uint16_t)v;
}

const uint64_t *doudizhu_rows_on_device(int device) {
    return (device >= 0 && device < 64 && g_dev_blob[device]) ? g_tab[device].rows : nullptr;
}

cudaError_t doudizhu_upload(int device, const void *blob, size_t nbytes) {
    if (device < 0 || device >= 64 || !blob || nbytes < sizeof(DdzBlobHeader)) return cudaErrorInvalidValue;
    DdzBlobHeader h; memcpy(&h, blob, sizeof h);
    if (memcmp(h.magic, "DDZ2", 4) != 0 || h.n_actions != 27472 || h.n_words_padded != 896 || h.total != nbytes || h.n_types != 38) return cudaErrorInvalidValue;
    // every array the kernels index must lie inside the blob (and be aligned for its element type)
    const auto inside = [&](uint64_t off, uint64_t bytes, uint64_t align) { return off % align == 0 && off <= nbytes && bytes <= nbytes - off; };
    if (!inside(h.off_rows, 8ull * h.n_actions, 8) || !inside(h.off_need, 16ull * h.n_words_padded, 16) ||
        !inside(h.off_type, h.n_actions, 1) || !inside(h.off_weight, h.n_actions, 1) || !inside(h.off_tw, 4ull * 38 * 17, 4))
        return cudaErrorInvalidValue;
    int prev = 0; cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    if (g_dev_blob[device]) { cudaFree(g_dev_blob[device]); g_dev_blob[device] = nullptr; memset(&g_tab[device], 0, sizeof g_tab[device]); }
    if (g_therm[device]) { cudaFree(g_therm[device]); g_therm[device] = nullptr; }
    void *dev_blob = nullptr;
    uint16_t *therm = nullptr;
    e = cudaMalloc(&dev_blob, nbytes);
    if (e == cudaSuccess) e = cudaMemcpy(dev_blob, blob, nbytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&therm, (size_t)h.n_actions * 27 * sizeof(uint16_t));
    if (e == cudaSuccess) {
        const int units = (int)h.n_actions * 27;
        k_ddz_build_therm<<<(units + 255) / 256, 256>>>(reinterpret_cast<const uint64_t *>(reinterpret_cast<const char *>(dev_blob) + h.off_rows), (int)h.n_actions, therm);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    if (e != cudaSuccess) {                                   // nothing half-installed: later calls keep answering NotReady
        if (dev_blob) cudaFree(dev_blob);
        if (therm) cudaFree(therm);
        cudaSetDevice(prev);
        return e;
    }
    g_dev_blob[device] = dev_blob;
    g_therm[device] = therm;
    {
        const char *b = reinterpret_cast<const char *>(g_dev_blob[device]);
        g_tab[device].rows = reinterpret_cast<const uint64_t *>(b + h.off_rows);
        g_tab[device].need = reinterpret_cast<const ulonglong2 *>(b + h.off_need);
        g_tab[device].type = reinterpret_cast<const uint8_t *>(b + h.off_type);
        g_tab[device].weight = reinterpret_cast<const uint8_t *>(b + h.off_weight);
        g_tab[device].tw_start = reinterpret_cast<const uint32_t *>(b + h.off_tw);
        g_tab[device].therm = therm;
    }
    cudaSetDevice(prev);
    return e;
}

// judger.py:124-258 playable_cards_from_hand (target < 0) / utils.py:225-262 get_gt_cards as a standalone
// operator: one warp per case, bit-packed legal set out
__global__ void __launch_bounds__(128) k_judge_doudizhu(const uint8_t *hands, const int32_t *targets, int n, uint32_t *mask, KParams p) {
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int i = blockIdx.x * 4 + wib;
    if (i >= n) return;
    constexpr int kWarpBytes = ((Doudizhu::MASK_WORDS * 4 + 15) & ~15) + Doudizhu::kScratchBytes;
    uint8_t *base = reinterpret_cast<uint8_t *>(smem_raw) + (size_t)wib * kWarpBytes;
    uint32_t *smask = reinterpret_cast<uint32_t *>(base);
    uint8_t *scratch = base + ((Doudizhu::MASK_WORDS * 4 + 15) & ~15);
    for (int w = lane; w < Doudizhu::MASK_WORDS; w += 32) smask[w] = 0;
    __syncwarp();
    Doudizhu g; g.bind(p, scratch);
    uint64_t h = 0;
    for (int r = 0; r < 15; r++) h |= (uint64_t)(hands[(size_t)i * 15 + r] & 15) << (4 * r);
    g.hand[0] = h; g.hand[1] = g.hand[2] = 0; g.played[0] = g.played[1] = g.played[2] = 0;
    g.cur = 0; g.winner = 3;
    const int t = targets ? targets[i] : -1;
    g.greater = t < 0 ? 3 : 1; g.greater_action = t < 0 ? (uint32_t)kDdzPass : (uint32_t)t;
    g.legal(smask, scratch, lane);
    __syncwarp();
    for (int w = lane; w < Doudizhu::MASK_WORDS; w += 32) mask[(size_t)i * Doudizhu::MASK_WORDS + w] = smask[w];
}
cudaError_t judge_doudizhu(const uint8_t *hands, const int32_t *targets, int n, uint32_t *mask, cudaStream_t s) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !g_dev_blob[dev]) return cudaErrorNotReady;
    KParams q; memset(&q, 0, sizeof q);
    DdzTables t = g_tab[dev];
    memcpy(q.tab, &t, sizeof t);
    const size_t smem = 4 * (size_t)(((Doudizhu::MASK_WORDS * 4 + 15) & ~15) + Doudizhu::kScratchBytes);
    k_judge_doudizhu<<<(n + 3) / 4, 128, smem, s>>>(hands, targets, n, mask, q);
    return cudaGetLastError();
}

cudaError_t dispatch_doudizhu(int op, int chance, int obs_dtype, const KParams &p, cudaStream_t s) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !g_dev_blob[dev]) return cudaErrorNotReady;       // rlc_upload_tables() first
    KParams q = p;
    q.tables = &g_tab[dev];
    DdzTables t = g_tab[dev];
    memcpy(q.tab, &t, sizeof t);
    if (obs_dtype == RLC_U8) return dispatch_wgame<Doudizhu, uint8_t>(op, chance, q, s);
    if (obs_dtype == RLC_F32) return dispatch_wgame<Doudizhu, float>(op, chance, q, s);
    return cudaErrorInvalidValue;
}
}  // namespace rlc
