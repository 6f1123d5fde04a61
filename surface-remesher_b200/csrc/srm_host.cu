// srm_host.cu — host side of the drop-in boundary: moving the caller's PAGEABLE buffers (main.cpp:203,214-215 hands
// malloc memory) to and from the device at PCIe speed, and reading the two sparse inputs on the host instead of
// uploading them.
//
// The reference does three blocking cudaMemcpy of pageable memory per gCVT call (gcvt.cu:906-914, 1101-1103) and one
// back (gcvt.cu:1152); a pageable cudaMemcpy is staged by the driver through one small pinned buffer and reaches
// 10-12 GB/s on this box, so the 872 MB of a 8192^2 call cost ~70 ms next to 4 ms of Lloyd iterations.  Here:
//   * srm_h2d_pageable / srm_d2h_pageable: T worker threads, each with two pinned 4 MB buffers and its own stream,
//     copy alternate chunks (memcpy pageable <-> pinned overlapped with the DMA of the previous chunk);
//   * srm_scan_site_map: the seed map (4 B/px, a handful of sites) is scanned by the threads into the packed site list
//     (row-major order, the order the device compaction produces), 400 KB go up instead of 268 MB;
//   * srm_scan_mask: the constraint mask (1 B/px) likewise into a list of pixels.
#include "srm_common.cuh"

#include <string.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

namespace {

constexpr int MAX_T = 16;
size_t CHUNK = 4u << 20;   // staging buffer size; srm_host_config() may change it (the pool is rebuilt)
int g_threads = 0;         // 0 = default: SRM_HOST_THREADS or half the hardware threads, at most 8

struct Lane {   // per worker thread: two pinned staging buffers, a stream, an event per buffer
    void *pin[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
};
struct Pool {
    int device = -1, T = 0;
    Lane lane[MAX_T];
};
std::mutex g_mu;
Pool g_pool;

int pool_threads() {
    static const int dflt = []() {
        const char *e = getenv("SRM_HOST_THREADS");
        int v = e ? atoi(e) : std::min(8, (int)std::thread::hardware_concurrency() / 2);
        return std::max(1, std::min(MAX_T, v));
    }();
    return g_threads > 0 ? g_threads : dflt;
}

// (re)creates the staging buffers for the current device; returns cudaSuccess or the first error
cudaError_t pool_get(Pool **out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (g_pool.device != dev) {
        for (int t = 0; t < g_pool.T; ++t) {
            Lane &l = g_pool.lane[t];
            for (int b = 0; b < 2; ++b) { if (l.pin[b]) cudaFreeHost(l.pin[b]); if (l.ev[b]) cudaEventDestroy(l.ev[b]); l.pin[b] = nullptr; l.ev[b] = nullptr; }
            if (l.st) cudaStreamDestroy(l.st);
            l.st = nullptr;
        }
        g_pool.T = 0;
        const int T = pool_threads();
        for (int t = 0; t < T; ++t) {
            Lane &l = g_pool.lane[t];
            for (int b = 0; b < 2; ++b) {
                if ((e = cudaHostAlloc(&l.pin[b], CHUNK, cudaHostAllocDefault)) != cudaSuccess) return e;
                if ((e = cudaEventCreateWithFlags(&l.ev[b], cudaEventDisableTiming)) != cudaSuccess) return e;
            }
            if ((e = cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking)) != cudaSuccess) return e;
            g_pool.T = t + 1;
        }
        g_pool.device = dev;
    }
    *out = &g_pool;
    return cudaSuccess;
}

template <typename F>
void parallel(int T, F f) {
    std::vector<std::thread> th;
    th.reserve(T > 0 ? T - 1 : 0);
    for (int t = 1; t < T; ++t) th.emplace_back(f, t);
    f(0);
    for (auto &x : th) x.join();
}

}  // namespace

// Measurement / tuning: worker threads (1..16, 0 = default) and staging chunk size in KB (256..65536, 0 = keep) of the
// pageable-copy pipeline and the host scans.  The staging pool is rebuilt on the next copy.
void srm_host_pool_release();
int srm_host_set_config(int threads, int chunk_kb) {
    if (threads < 0 || threads > MAX_T || (chunk_kb != 0 && (chunk_kb < 256 || chunk_kb > 65536))) return -1;
    srm_host_pool_release();
    std::lock_guard<std::mutex> lock(g_mu);
    g_threads = threads;
    if (chunk_kb) CHUNK = (size_t)chunk_kb << 10;
    return 0;
}

void srm_host_pool_release() {
    std::lock_guard<std::mutex> lock(g_mu);
    for (int t = 0; t < g_pool.T; ++t) {
        Lane &l = g_pool.lane[t];
        for (int b = 0; b < 2; ++b) { if (l.pin[b]) cudaFreeHost(l.pin[b]); if (l.ev[b]) cudaEventDestroy(l.ev[b]); l.pin[b] = nullptr; l.ev[b] = nullptr; }
        if (l.st) cudaStreamDestroy(l.st);
        l.st = nullptr;
    }
    g_pool.T = 0; g_pool.device = -1;
}

// Blocking copy of `bytes` from pageable (or any) host memory to device memory.  `after`: work already enqueued on this
// stream that reads/writes dst must finish first (an event on it gates the copy streams).
cudaError_t srm_h2d_pageable(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t after) {
    if (bytes == 0) return cudaSuccess;
    std::lock_guard<std::mutex> lock(g_mu);
    Pool *P = nullptr;
    cudaError_t e = pool_get(&P);
    if (e != cudaSuccess) return e;
    const int dev = P->device;
    const size_t nchunk = (bytes + CHUNK - 1) / CHUNK;
    const int T = (int)std::min<size_t>((size_t)P->T, nchunk);
    cudaEvent_t gate = nullptr;
    if (after) {
        if ((e = cudaEventCreateWithFlags(&gate, cudaEventDisableTiming)) != cudaSuccess) return e;
        cudaEventRecord(gate, after);
    }
    std::vector<cudaError_t> err((size_t)T, cudaSuccess);
    parallel(T, [&](int t) {
        cudaSetDevice(dev);
        Lane &l = P->lane[t];
        cudaError_t le = cudaSuccess;
        if (gate) le = cudaStreamWaitEvent(l.st, gate, 0);
        int k = 0;
        for (size_t c = (size_t)t; c < nchunk && le == cudaSuccess; c += (size_t)T, ++k) {
            const int b = k & 1;
            const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
            if (k >= 2) le = cudaEventSynchronize(l.ev[b]);   // the DMA that last read this buffer is done
            if (le != cudaSuccess) break;
            memcpy(l.pin[b], (const char *)src_host + off, len);
            le = cudaMemcpyAsync((char *)dst_dev + off, l.pin[b], len, cudaMemcpyHostToDevice, l.st);
            if (le == cudaSuccess) le = cudaEventRecord(l.ev[b], l.st);
        }
        if (le == cudaSuccess) le = cudaStreamSynchronize(l.st);
        err[(size_t)t] = le;
    });
    if (gate) cudaEventDestroy(gate);
    for (cudaError_t x : err) if (x != cudaSuccess) return x;
    return cudaSuccess;
}

// Blocking copy device -> pageable host memory; work enqueued on `after` is waited for first.
cudaError_t srm_d2h_pageable(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t after) {
    if (bytes == 0) return cudaSuccess;
    std::lock_guard<std::mutex> lock(g_mu);
    Pool *P = nullptr;
    cudaError_t e = pool_get(&P);
    if (e != cudaSuccess) return e;
    const int dev = P->device;
    const size_t nchunk = (bytes + CHUNK - 1) / CHUNK;
    const int T = (int)std::min<size_t>((size_t)P->T, nchunk);
    cudaEvent_t gate = nullptr;
    if (after) {
        if ((e = cudaEventCreateWithFlags(&gate, cudaEventDisableTiming)) != cudaSuccess) return e;
        cudaEventRecord(gate, after);
    }
    std::vector<cudaError_t> err((size_t)T, cudaSuccess);
    parallel(T, [&](int t) {
        cudaSetDevice(dev);
        Lane &l = P->lane[t];
        cudaError_t le = cudaSuccess;
        if (gate) le = cudaStreamWaitEvent(l.st, gate, 0);
        // chunks of this lane: c_k = t + k T.  Keep one DMA in flight while the previous chunk is copied out.
        auto issue = [&](int k) -> cudaError_t {
            const size_t c = (size_t)t + (size_t)k * (size_t)T;
            if (c >= nchunk) return cudaSuccess;
            const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
            cudaError_t x = cudaMemcpyAsync(l.pin[k & 1], (const char *)src_dev + off, len, cudaMemcpyDeviceToHost, l.st);
            return x == cudaSuccess ? cudaEventRecord(l.ev[k & 1], l.st) : x;
        };
        if (le == cudaSuccess) le = issue(0);
        for (int k = 0; le == cudaSuccess; ++k) {
            const size_t c = (size_t)t + (size_t)k * (size_t)T;
            if (c >= nchunk) break;
            le = issue(k + 1);
            if (le != cudaSuccess) break;
            le = cudaEventSynchronize(l.ev[k & 1]);
            if (le != cudaSuccess) break;
            const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
            memcpy((char *)dst_host + off, l.pin[k & 1], len);
        }
        err[(size_t)t] = le;
    });
    if (gate) cudaEventDestroy(gate);
    for (cudaError_t x : err) if (x != cudaSuccess) return x;
    return cudaSuccess;
}

// Sites of a dense seed map (pixels whose x half is not MARKER, gcvt.cu:90,:240) as a packed list in row-major scan
// order — the order the device compaction (k_sites_count / k_sites_write) produces, so site ids do not depend on the path.
// The map is almost empty (K sites in N pixels), so the scan is a stream over 4 B/px: blocks of 64 pixels are tested
// with 128-bit words (x halves against the marker, OR-reduced; 64-bit words without SSE2) and only blocks that hold a
// site are looked at pixel by pixel.  (The per-pixel loop this replaces took 15-35 ms at 8192^2 on 4-8
// threads — longer than the density upload it is meant to hide behind.)
void srm_scan_site_map(const int *site_map, size_t N, std::vector<int> &sites) {
    const int T = pool_threads();
    std::vector<std::vector<int>> part((size_t)T);
    parallel(T, [&](int t) {
        const size_t i0 = N * (size_t)t / (size_t)T, i1 = N * (size_t)(t + 1) / (size_t)T;
        std::vector<int> &out = part[(size_t)t];
        const unsigned char *base = reinterpret_cast<const unsigned char *>(site_map);   // short-aligned at least
        constexpr unsigned long long XH = 0x0000ffff0000ffffull, MK = 0x0000800000008000ull;   // little endian: x | y << 16
        auto pixel = [&](size_t i) {
            int p;
            memcpy(&p, base + 4 * i, 4);
            if ((short)(p & 0xffff) != (short)SRM_MARK) out.push_back(p);
        };
        size_t i = i0;
#if defined(__SSE2__)
        const __m128i mk = _mm_set1_epi32(0x00008000), xh = _mm_set1_epi32(0x0000ffff), zero = _mm_setzero_si128();
        for (; i + 64 <= i1; i += 64) {   // 256 bytes per step, four independent OR chains: runs at the thread's read rate
            const __m128i *w = reinterpret_cast<const __m128i *>(base + 4 * i);
            __m128i a0 = zero, a1 = zero, a2 = zero, a3 = zero;
            for (int q = 0; q < 16; q += 4) {
                a0 = _mm_or_si128(a0, _mm_xor_si128(_mm_loadu_si128(w + q), mk));
                a1 = _mm_or_si128(a1, _mm_xor_si128(_mm_loadu_si128(w + q + 1), mk));
                a2 = _mm_or_si128(a2, _mm_xor_si128(_mm_loadu_si128(w + q + 2), mk));
                a3 = _mm_or_si128(a3, _mm_xor_si128(_mm_loadu_si128(w + q + 3), mk));
            }
            const __m128i any = _mm_and_si128(_mm_or_si128(_mm_or_si128(a0, a1), _mm_or_si128(a2, a3)), xh);
            if (_mm_movemask_epi8(_mm_cmpeq_epi8(any, zero)) == 0xffff) continue;
            for (size_t q = 0; q < 64; ++q) pixel(i + q);
        }
#endif
        for (; i + 32 <= i1; i += 32) {
            unsigned long long w[16], any = 0;
            memcpy(w, base + 4 * i, sizeof(w));
            for (int q = 0; q < 16; ++q) any |= (w[q] ^ MK) & XH;
            if (!any) continue;
            for (size_t q = 0; q < 32; ++q) pixel(i + q);
        }
        for (; i < i1; ++i) pixel(i);
    });
    size_t tot = 0;
    for (auto &v : part) tot += v.size();
    sites.clear();
    sites.reserve(tot);
    for (auto &v : part) sites.insert(sites.end(), v.begin(), v.end());
}

// Non-zero bytes of the constraint mask as packed pixels (x | y << 16), any order.
void srm_scan_mask(const unsigned char *mask, int n, std::vector<int> &pixels, int row0, int row1) {
    const int T = pool_threads();
    if (row1 < 0) row1 = n;
    std::vector<std::vector<int>> part((size_t)T);
    parallel(T, [&](int t) {
        const int y0 = row0 + (int)((long long)(row1 - row0) * t / T), y1 = row0 + (int)((long long)(row1 - row0) * (t + 1) / T);
        std::vector<int> &out = part[(size_t)t];
        for (int y = y0; y < y1; ++y) {
            const unsigned char *row = mask + (size_t)y * n;
            int x0 = 0;
            for (; x0 + 64 <= n; x0 += 64) {   // blocks of 64 pixels, OR-reduced first: the mask is almost empty
                unsigned long long w[8], any = 0;
                memcpy(w, row + x0, sizeof(w));
                for (int q = 0; q < 8; ++q) any |= w[q];
                if (!any) continue;
                for (int b = 0; b < 64; ++b)
                    if (row[x0 + b]) out.push_back(srm_pack(x0 + b, y));
            }
            for (; x0 < n; ++x0)
                if (row[x0]) out.push_back(srm_pack(x0, y));
        }
    });
    pixels.clear();
    for (auto &v : part) pixels.insert(pixels.end(), v.begin(), v.end());
}
