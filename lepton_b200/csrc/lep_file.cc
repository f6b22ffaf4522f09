// lep_file.cc -- file-level drop-in: JPEG bytes -> .lep bytes and back, batched.
//
// Host threads do what the reference's jpgcoder.cc does around the codec boundary (read_jpeg / decode_jpeg /
// write_ujpg on the way in, read_ujpg / recode_baseline_jpeg on the way out); the arithmetic coding itself goes
// through the C ABI of lep_capi.cu to the sm_100a kernels.  No CPU coder exists in this library.
#include <atomic>
#include <cstring>
#include <memory>
#include <thread>
#include <array>
#include <algorithm>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/lepton_b200.h"
#include "lep_host.h"

using namespace lephost;

struct lepb200_codec {
    lepb200_ctx* ctx = nullptr;
    int nthreads = 1;
    void* arena = nullptr;          // pinned host memory for coefficient planes
    size_t arena_cap = 0;
    std::vector<std::vector<uint8_t>> outputs;
    std::string err;
    // timing of the last call (seconds): parse+huffman, gpu (upload+kernel+fetch), container
    double t_front = 0, t_gpu = 0, t_back = 0;
};

namespace {

template <class F>
void parallel_for(int n, int nthreads, F&& f) {
    nthreads = std::max(1, std::min(nthreads, n));
    if (nthreads == 1) { for (int i = 0; i < n; ++i) f(i); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&]() { for (int i; (i = next.fetch_add(1)) < n;) f(i); });
    for (auto& t : th) t.join();
}

double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

bool reserve_arena(lepb200_codec* c, size_t bytes) {
    if (bytes <= c->arena_cap) return true;
    if (c->arena) lepb200_pinned_free(c->arena);
    c->arena = nullptr; c->arena_cap = 0;
    size_t want = bytes + bytes / 8 + 4096;
    c->arena = lepb200_pinned_alloc(want);
    if (!c->arena) return false;
    c->arena_cap = want;
    return true;
}

void fill_image(lepb200_image& im, const Jpeg& j, int16_t* const planes[4], const std::vector<Handoff>& sel) {
    memset(&im, 0, sizeof(im));
    im.ncmp = j.ncmp; im.mcuv = j.mcuv;
    for (int c = 0; c < j.ncmp; ++c) {
        im.bch[c] = j.cmp[c].bch; im.bcv[c] = j.cmp[c].bcv;
        im.trunc_bcv[c] = j.cmp[c].bcv; im.trunc_bc[c] = j.cmp[c].bc;
        memcpy(im.qtable_zigzag[c], j.qtables[j.cmp[c].tq], 128);
        im.planes[c] = planes[c];
    }
    im.nseg = (int)sel.size();
    for (size_t s = 0; s < sel.size(); ++s) im.luma_y_start[s] = sel[s].luma_y_start;
}

}  // namespace

extern "C" {

int lepb200_codec_create(lepb200_codec** out, int device, int host_threads) {
    if (!out) return LEPB200_ERR_INVALID;
    *out = nullptr;
    lepb200_ctx* ctx = nullptr;
    int r = lepb200_create(&ctx, device);
    if (r) return r;
    lepb200_codec* c = new lepb200_codec();
    c->ctx = ctx;
    c->nthreads = host_threads > 0 ? host_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    *out = c;
    return LEPB200_OK;
}

void lepb200_codec_destroy(lepb200_codec* c) {
    if (!c) return;
    if (c->arena) lepb200_pinned_free(c->arena);
    lepb200_destroy(c->ctx);
    delete c;
}

const char* lepb200_codec_last_error(const lepb200_codec* c) {
    if (!c) return "null codec";
    return c->err.empty() ? lepb200_last_error(c->ctx) : c->err.c_str();
}

lepb200_ctx* lepb200_codec_ctx(lepb200_codec* c) { return c ? c->ctx : nullptr; }

void lepb200_codec_last_timing(const lepb200_codec* c, double* front_s, double* gpu_s, double* back_s) {
    if (!c) return;
    if (front_s) *front_s = c->t_front;
    if (gpu_s) *gpu_s = c->t_gpu;
    if (back_s) *back_s = c->t_back;
}

// JPEG files -> .lep files.  out[i].data points into codec-owned memory, valid until the next call.
int lepb200_compress_jpegs(lepb200_codec* c, const lepb200_buffer* jpegs, int n, lepb200_result* out) {
    if (!c || !jpegs || !out || n <= 0) return LEPB200_ERR_INVALID;
    c->err.clear();
    double t0 = now_s();
    std::vector<std::unique_ptr<Jpeg>> js(n);
    // plane arena layout from a header-only peek (so that the big per-image buffers are never malloc'ed)
    std::vector<size_t> base(n, 0), need(n, 0);
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
        need[i] = peek_plane_bytes(jpegs[i].data, jpegs[i].len);
        base[i] = total;
        total += need[i];
    }
    if (!reserve_arena(c, total + 256)) { c->err = "pinned host allocation failed"; return LEPB200_ERR_NOMEM; }
    std::vector<std::array<int16_t*, 4>> planes(n);
    std::vector<Splits> splits(n);
    parallel_for(n, c->nthreads, [&](int i) {
        // the de-stuffed entropy buffer is the only large per-image allocation: recycle it per thread
        static thread_local std::vector<uint8_t> huff_scratch;
        static thread_local std::vector<std::pair<uint32_t, uint32_t>> offs_scratch;
        js[i].reset(new Jpeg());
        Jpeg& j = *js[i];
        huff_scratch.clear(); offs_scratch.clear();
        j.huff.swap(huff_scratch); j.offs.swap(offs_scratch);
        bool ok = parse_jpeg(jpegs[i].data, jpegs[i].len, j);
        if (ok) {
            size_t want = 0;
            for (int k = 0; k < j.ncmp; ++k) want += (plane_bytes(j, k) + 255) & ~size_t(255);
            if (want != need[i]) { j.status = NOT_HANDLED; j.error = "plane size peek mismatch"; ok = false; }
        }
        if (ok) {
            uint8_t* p = static_cast<uint8_t*>(c->arena) + base[i];
            for (int k = 0; k < 4; ++k) planes[i][k] = nullptr;
            for (int k = 0; k < j.ncmp; ++k) {
                planes[i][k] = reinterpret_cast<int16_t*>(p);
                size_t pb = plane_bytes(j, k);
                memset(p, 0, pb);
                p += (pb + 255) & ~size_t(255);
            }
            if (decode_scans(j, planes[i].data())) splits[i] = select_splits(j);
        }
        j.huff.swap(huff_scratch); j.offs.swap(offs_scratch);     // keep the capacity with the thread
    });
    double t1 = now_s();
    // GPU: one batch over all images that survived the front end
    std::vector<lepb200_image> imgs;
    std::vector<int> idx;
    int nseg_total = 0;
    for (int i = 0; i < n; ++i) {
        if (js[i]->status) continue;
        lepb200_image im;
        fill_image(im, *js[i], planes[i].data(), splits[i].selected);
        imgs.push_back(im);
        idx.push_back(i);
        nseg_total += im.nseg;
    }
    std::vector<lepb200_stream> streams(nseg_total);
    if (!imgs.empty()) {
        int r = lepb200_encode_images(c->ctx, imgs.data(), (int)imgs.size(), streams.data());
        if (r) return r;
    }
    double t2 = now_s();
    c->outputs.assign(n, std::vector<uint8_t>());
    std::vector<int> seg_base(imgs.size() + 1, 0);
    for (size_t k = 0; k < imgs.size(); ++k) seg_base[k + 1] = seg_base[k] + imgs[k].nseg;
    std::vector<int> status(n, 0);
    for (int i = 0; i < n; ++i) status[i] = js[i]->status;
    parallel_for((int)imgs.size(), c->nthreads, [&](int k) {
        const int i = idx[k];
        std::vector<std::pair<const uint8_t*, size_t>> ss;
        for (int s = seg_base[k]; s < seg_base[k + 1]; ++s) {
            if (streams[s].status) { status[i] = streams[s].status; return; }
            ss.emplace_back(streams[s].data, (size_t)streams[s].len);
        }
        std::string err;
        if (!write_lep(*js[i], splits[i], ss, c->outputs[i], err)) { status[i] = NOT_HANDLED; c->outputs[i].clear(); }
    });
    double t3 = now_s();
    for (int i = 0; i < n; ++i) {
        out[i].status = status[i];
        out[i].data = status[i] ? nullptr : c->outputs[i].data();
        out[i].len = status[i] ? 0 : c->outputs[i].size();
    }
    c->t_front = t1 - t0; c->t_gpu = t2 - t1; c->t_back = t3 - t2;
    return LEPB200_OK;
}

// Host front end only (parse + Huffman decode + split selection) over a batch with `threads` workers; returns the
// wall-clock seconds.  Diagnostic: lets the host stage be profiled without a GPU.
double lepb200_host_frontend_seconds(const lepb200_buffer* jpegs, int n, int threads, int32_t* first_error) {
    std::vector<std::unique_ptr<Jpeg>> js(n);
    std::vector<std::vector<int16_t>> store(n);
    double t0 = now_s();
    parallel_for(n, threads, [&](int i) {
        js[i].reset(new Jpeg());
        Jpeg& j = *js[i];
        if (!parse_jpeg(jpegs[i].data, jpegs[i].len, j)) return;
        size_t total = 0;
        for (int k = 0; k < j.ncmp; ++k) total += (size_t)j.cmp[k].bc * 64;
        store[i].assign(total, 0);
        int16_t* planes[4] = {nullptr, nullptr, nullptr, nullptr};
        size_t off = 0;
        for (int k = 0; k < j.ncmp; ++k) { planes[k] = store[i].data() + off; off += (size_t)j.cmp[k].bc * 64; }
        if (decode_scans(j, planes)) select_splits(j);
    });
    double dt = now_s() - t0;
    if (first_error) { *first_error = 0; for (int i = 0; i < n; ++i) if (js[i]->status) { *first_error = js[i]->status; break; } }
    return dt;
}

// ---- staged host-only entry points (no GPU involved): parse + Huffman-decode one JPEG, expose its planes and
// thread-segment split as a lepb200_image, and assemble the .lep from externally coded segment streams.
struct lepb200_jpeg {
    Jpeg j;
    Splits sp;
    std::vector<std::vector<int16_t>> store;
    int16_t* planes[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<uint8_t> out;
};

int lepb200_host_jpeg_open(const uint8_t* data, size_t len, lepb200_jpeg** out, int32_t* status) {
    if (!out || !data) return LEPB200_ERR_INVALID;
    lepb200_jpeg* h = new lepb200_jpeg();
    *out = h;
    if (parse_jpeg(data, len, h->j)) {
        h->store.resize(h->j.ncmp);
        for (int c = 0; c < h->j.ncmp; ++c) {
            h->store[c].assign((size_t)h->j.cmp[c].bc * 64, 0);
            h->planes[c] = h->store[c].data();
        }
        if (decode_scans(h->j, h->planes)) h->sp = select_splits(h->j);
    }
    if (status) *status = h->j.status;
    return LEPB200_OK;
}

const char* lepb200_host_jpeg_error(const lepb200_jpeg* h) { return h ? h->j.error.c_str() : "null"; }

int lepb200_host_jpeg_image(lepb200_jpeg* h, lepb200_image* img) {
    if (!h || !img || h->j.status) return LEPB200_ERR_INVALID;
    fill_image(*img, h->j, h->planes, h->sp.selected);
    return LEPB200_OK;
}

int lepb200_host_jpeg_write_lep(lepb200_jpeg* h, const lepb200_stream* streams, int nseg, const uint8_t** data, size_t* len) {
    if (!h || !streams || !data || !len || h->j.status || nseg != (int)h->sp.selected.size()) return LEPB200_ERR_INVALID;
    std::vector<std::pair<const uint8_t*, size_t>> ss;
    for (int s = 0; s < nseg; ++s) ss.emplace_back(streams[s].data, (size_t)streams[s].len);
    std::string err;
    if (!write_lep(h->j, h->sp, ss, h->out, err)) { h->j.error = err; return LEPB200_ERR_INVALID; }
    *data = h->out.data();
    *len = h->out.size();
    return LEPB200_OK;
}

void lepb200_host_jpeg_close(lepb200_jpeg* h) { delete h; }

}  // extern "C"
