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

cudaError_t doudizhu_upload(int device, const void *blob, size_t nbytes) {
    if (device < 0 || device >= 64 || !blob || nbytes < sizeof(DdzBlobHeader)) return cudaErrorInvalidValue;
    DdzBlobHeader h; memcpy(&h, blob, sizeof h);
    if (memcmp(h.magic, "DDZ1", 4) != 0 || h.n_actions != 27472 || h.n_words_padded != 896 || h.total != nbytes || h.n_types != 38) return cudaErrorInvalidValue;
    int prev = 0; cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    if (g_dev_blob[device]) { cudaFree(g_dev_blob[device]); g_dev_blob[device] = nullptr; }
    e = cudaMalloc(&g_dev_blob[device], nbytes);
    if (e == cudaSuccess) e = cudaMemcpy(g_dev_blob[device], blob, nbytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const char *b = reinterpret_cast<const char *>(g_dev_blob[device]);
        g_tab[device].rows = reinterpret_cast<const uint64_t *>(b + h.off_rows);
        g_tab[device].need = reinterpret_cast<const uint64_t *>(b + h.off_need);
        g_tab[device].type = reinterpret_cast<const uint8_t *>(b + h.off_type);
        g_tab[device].weight = reinterpret_cast<const uint8_t *>(b + h.off_weight);
        g_tab[device].tw_start = reinterpret_cast<const uint32_t *>(b + h.off_tw);
    }
    cudaSetDevice(prev);
    return e;
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
