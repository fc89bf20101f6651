// mgpu.cu -- multi-GPU entry points BELOW the C ABI, so that a caller that is not
// Python + torch.distributed (the Haskell shim: one process, one library call) fans out over the GPUs of
// the box the way the reference's ...P functions fan out over the cores (parListChunk over
// getNumCapabilities, src/Data/FMIndex.hs:417-423, 544-553):
//   tc_mgpu_blocks_encode_packed  independent blocks round-robin over the devices, no exchange step
//   tc_fm_replicate               the index image copied device to device (peer copies over NVLink)
//   tc_mgpu_fm_count / _locate    contiguous query chunks per device, results in input order
// One host thread per device; each device has ONE pooled context (created on first use, kept for the life of
// the process, guarded by a mutex), which is also what tc_ctx_pool_acquire hands to single-GPU callers that
// do not want to create and destroy a context (stream + arena + pinned page) per call.
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"
#include "impl.cuh"

namespace {
struct PoolEntry {
    std::mutex busy; // held while a call runs on the context
    tc_ctx *ctx = nullptr;
};
std::mutex g_pool_mu;
std::vector<PoolEntry *> g_pool; // index = device

PoolEntry *pool_entry(int device) {
    std::lock_guard<std::mutex> g(g_pool_mu);
    if (device < 0) return nullptr;
    if ((size_t)device >= g_pool.size()) g_pool.resize(device + 1, nullptr);
    if (!g_pool[device]) g_pool[device] = new PoolEntry();
    return g_pool[device];
}

// Runs fn(ctx of devices[i], i) on one thread per device; the first error (other than TC_E_CAP) wins.
template <class F>
int per_device(int ndev, const int *devices, F fn) {
    if (ndev <= 0 || !devices) return TC_E_ARG;
    std::vector<int> rc(ndev, TC_OK);
    std::vector<std::thread> th;
    auto body = [&](int i) {
        PoolEntry *e = pool_entry(devices[i]);
        if (!e) {
            rc[i] = TC_E_ARG;
            return;
        }
        std::lock_guard<std::mutex> g(e->busy);
        if (!e->ctx) {
            int r = tc_ctx_create(devices[i], &e->ctx);
            if (r != TC_OK) {
                rc[i] = r;
                return;
            }
        }
        cudaSetDevice(devices[i]);
        rc[i] = fn(e->ctx, i);
    };
    for (int i = 1; i < ndev; i++) th.emplace_back(body, i);
    body(0);
    for (auto &t : th) t.join();
    int out = TC_OK;
    for (int i = 0; i < ndev; i++) {
        if (rc[i] == TC_OK) continue;
        if (rc[i] != TC_E_CAP) return rc[i];
        out = TC_E_CAP;
    }
    return out;
}
} // namespace

extern "C" int tc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int tc_ctx_pool_acquire(int device, tc_ctx **out) {
    if (!out) return TC_E_ARG;
    *out = nullptr;
    PoolEntry *e = pool_entry(device);
    if (!e) return TC_E_ARG;
    e->busy.lock();
    if (!e->ctx) {
        int r = tc_ctx_create(device, &e->ctx);
        if (r != TC_OK) {
            e->busy.unlock();
            return r;
        }
    }
    cudaSetDevice(device);
    *out = e->ctx;
    return TC_OK;
}
extern "C" void tc_ctx_pool_release(tc_ctx *ctx) {
    if (!ctx) return;
    PoolEntry *e = pool_entry(ctx->device);
    if (e && e->ctx == ctx) e->busy.unlock();
}

extern "C" int tc_mgpu_blocks_encode_packed(int ndev, const int *devices, uint64_t nblocks, const uint8_t *const *text,
                                            const uint64_t *n, int with_mtf, uint8_t *const *out, const uint64_t *cap,
                                            uint64_t *out_bytes, tc_block_info *info) {
    if (nblocks == 0) return TC_OK;
    if (!text || !n || !out || !cap || !out_bytes || !info) return TC_E_ARG;
    return per_device(ndev, devices, [&](tc_ctx *ctx, int i) -> int {
        // block b belongs to device b mod ndev: gather this device's blocks into contiguous argument arrays
        std::vector<const uint8_t *> t;
        std::vector<uint64_t> nn, cc, ob;
        std::vector<uint8_t *> o;
        std::vector<tc_block_info> inf;
        for (uint64_t b = i; b < nblocks; b += ndev) {
            t.push_back(text[b]), nn.push_back(n[b]), o.push_back(out[b]), cc.push_back(cap[b]);
        }
        if (t.empty()) return TC_OK;
        ob.resize(t.size());
        inf.resize(t.size());
        int rc = tc_blocks_encode_packed(ctx, t.size(), t.data(), nn.data(), with_mtf, o.data(), cc.data(), ob.data(), inf.data());
        uint64_t k = 0;
        for (uint64_t b = i; b < nblocks; b += ndev, k++) out_bytes[b] = ob[k], info[b] = inf[k];
        return rc;
    });
}

extern "C" int tc_fm_replicate(const tc_fm *root, int ndev, const int *devices, tc_fm **replicas) {
    if (!root || !replicas) return TC_E_ARG;
    tc_fm_info ri;
    TC_TRY(tc_fm_get_info(root, &ri));
    const void *src = tc_fm_blob(root);
    int root_dev = -1;
    {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, src) != cudaSuccess) {
            cudaGetLastError();
            return TC_E_ARG;
        }
        root_dev = a.device;
    }
    for (int i = 0; i < ndev; i++) replicas[i] = nullptr;
    int rc = per_device(ndev, devices, [&](tc_ctx *ctx, int i) -> int {
        void *blob = nullptr;
        TC_CUDA(cudaMalloc(&blob, ri.blob_bytes));
        // device-to-device over NVLink when the two devices are peers; staged by the driver otherwise
        cudaError_t e = cudaMemcpyPeerAsync(blob, devices[i], src, root_dev, ri.blob_bytes, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            cudaFree(blob);
            return ctx->fail(e, "tc_fm_replicate", __LINE__);
        }
        int r = tc_fm_from_blob_dev(ctx, blob, ri.blob_bytes, 1, &replicas[i]);
        if (r != TC_OK) cudaFree(blob);
        return r;
    });
    if (rc != TC_OK)
        for (int i = 0; i < ndev; i++) {
            tc_fm_free(replicas[i]);
            replicas[i] = nullptr;
        }
    return rc;
}

namespace {
// [a, b) of q queries for device i of ndev: contiguous chunks like parListChunk (src/Data/FMIndex.hs:418-422)
inline void chunk_of(uint64_t q, int ndev, int i, uint64_t *a, uint64_t *b) {
    const uint64_t per = (q + ndev - 1) / ndev;
    *a = per * i < q ? per * i : q;
    *b = *a + per < q ? *a + per : q;
}
} // namespace

extern "C" int tc_mgpu_fm_count(int ndev, const int *devices, tc_fm *const *replicas, const uint8_t *pats,
                                const uint64_t *off, uint64_t q, int64_t *count) {
    if (q == 0) return TC_OK;
    if (!replicas || !pats || !off || !count) return TC_E_ARG;
    return per_device(ndev, devices, [&](tc_ctx *ctx, int i) -> int {
        uint64_t a, b;
        chunk_of(q, ndev, i, &a, &b);
        if (a == b) return TC_OK;
        std::vector<uint64_t> o(b - a + 1);
        for (uint64_t k = a; k <= b; k++) o[k - a] = off[k] - off[a];
        return tc_fm_count(ctx, replicas[i], pats + off[a], o.data(), b - a, count + a);
    });
}

extern "C" int tc_mgpu_fm_locate(int ndev, const int *devices, tc_fm *const *replicas, const uint8_t *pats,
                                 const uint64_t *off, uint64_t q, uint64_t *hit_off, uint64_t *pos_1based, uint64_t cap,
                                 uint64_t *total) {
    if (!total || !hit_off) return TC_E_ARG;
    *total = 0;
    hit_off[0] = 0;
    if (q == 0) return TC_OK;
    if (!replicas || !pats || !off) return TC_E_ARG;
    // pass 1: hits per pattern on every device (hit_off chunk-local for now), pass 2: positions at their final place
    std::vector<uint64_t> tot(ndev, 0);
    std::vector<std::vector<uint64_t>> offs(ndev);
    int rc = per_device(ndev, devices, [&](tc_ctx *ctx, int i) -> int {
        uint64_t a, b;
        chunk_of(q, ndev, i, &a, &b);
        if (a == b) return TC_OK;
        offs[i].resize(b - a + 1);
        for (uint64_t k = a; k <= b; k++) offs[i][k - a] = off[k] - off[a];
        std::vector<uint64_t> ho(b - a + 1);
        int r = tc_fm_locate(ctx, replicas[i], pats + off[a], offs[i].data(), b - a, ho.data(), nullptr, 0, &tot[i]);
        if (r != TC_OK && r != TC_E_CAP) return r;
        for (uint64_t k = a; k < b; k++) hit_off[k + 1] = ho[k - a + 1] - ho[k - a]; // per-pattern counts for now
        return TC_OK;
    });
    if (rc != TC_OK) return rc;
    for (uint64_t k = 0; k < q; k++) hit_off[k + 1] += hit_off[k];
    *total = hit_off[q];
    if (*total > cap || !pos_1based) return *total ? TC_E_CAP : TC_OK;
    return per_device(ndev, devices, [&](tc_ctx *ctx, int i) -> int {
        uint64_t a, b;
        chunk_of(q, ndev, i, &a, &b);
        if (a == b || tot[i] == 0) return TC_OK;
        std::vector<uint64_t> ho(b - a + 1);
        uint64_t t2 = 0;
        return tc_fm_locate(ctx, replicas[i], pats + off[a], offs[i].data(), b - a, ho.data(), pos_1based + hit_off[a], tot[i],
                            &t2);
    });
}
