// Multi-device context (pk_ctx_create_multi): ONE host thread calls the C ABI, N devices compute.
//
// The reference's zone scheme lives inside the multiplier: series_multiplier::operator()() cuts the output into zones,
// lets the threads of the pool claim and consume them, and returns one series (polynomial.hpp:1783-1966).  This is the
// same thing one level up: the output Kronecker-code range is cut into N ordered parts (balanced by work), device r
// computes part r from its own replica of the operands (a few MB each), and the parts leave the devices over N PCIe
// links at once, each straight to its offset of ONE pinned host result.  Ranges are ordered, so the host arrays are
// the sorted product without a merge; nothing travels between the devices for a host-to-host product.
//
// Device-resident products (pk_upload / pk_mul_dev / pk_download) stay partitioned in HBM -- part r on device r -- until
// the product is used as an operand; only then are the parts copied to every device (peer-to-peer over NVLink).
//
// Included by pk_abi.cu after the single-device implementation (upload_impl, mul_impl, pinned_acquire).
#pragma once

struct pk_multi_state {
    struct Worker {
        std::thread th;
        std::function<void()> job;
        bool has_job = false, done = true;
        int code = PK_OK;
        std::string msg;
    };
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    bool stop = false;
    std::vector<Worker> w;
};

namespace
{

struct DeviceScope { // makes `device` current for the calling thread and restores the caller's device afterwards
    int prev = -1;
    explicit DeviceScope(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) cudaSetDevice(device);
        else prev = -1;
    }
    ~DeviceScope()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceScope(const DeviceScope &) = delete;
    DeviceScope &operator=(const DeviceScope &) = delete;
};

void multi_worker_main(pk_ctx *M, u32 r)
{
    pk_multi_state &S = *M->multi;
    cudaSetDevice(M->subs[r]->device);
    std::unique_lock<std::mutex> lk(S.m);
    for (;;) {
        S.cv_job.wait(lk, [&] { return S.stop || S.w[r].has_job; });
        if (S.stop) return;
        std::function<void()> job = std::move(S.w[r].job);
        S.w[r].has_job = false;
        lk.unlock();
        int code = PK_OK;
        std::string msg;
        try {
            job();
        } catch (const PkError &e) {
            code = e.code;
            msg = e.msg;
        } catch (const std::bad_alloc &) {
            code = PK_E_NOMEM;
            msg = "host allocation failure";
        } catch (const std::exception &e) {
            code = PK_E_CUDA;
            msg = e.what();
        }
        lk.lock();
        S.w[r].code = code;
        S.w[r].msg = std::move(msg);
        S.w[r].done = true;
        S.cv_done.notify_all();
    }
}

// run fn(r) on the worker of every device, wait for all of them; the first failure (lowest device) is rethrown
void multi_run(pk_ctx *M, const std::function<void(u32)> &fn)
{
    pk_multi_state &S = *M->multi;
    const u32 N = (u32)M->subs.size();
    {
        std::unique_lock<std::mutex> lk(S.m);
        for (u32 r = 0; r < N; ++r) {
            S.w[r].job = [&fn, r] { fn(r); };
            S.w[r].has_job = true;
            S.w[r].done = false;
            S.w[r].code = PK_OK;
        }
        S.cv_job.notify_all();
        S.cv_done.wait(lk, [&] {
            for (u32 r = 0; r < N; ++r)
                if (!S.w[r].done) return false;
            return true;
        });
    }
    for (u32 r = 0; r < N; ++r)
        if (S.w[r].code != PK_OK) fail(S.w[r].code, "device " + std::to_string(M->subs[r]->device) + ": " + S.w[r].msg);
}

void free_single(pk_ctx *sub, pk_dpoly *p)
{
    if (!p) return;
    DeviceScope ds(sub->device);
    free_dpoly_arrays(sub, p);
    delete p;
}

void multi_free(pk_ctx *M, pk_dpoly *p)
{
    if (!p) return;
    for (size_t r = 0; r < p->rep.size() && r < M->subs.size(); ++r) free_single(M->subs[r], p->rep[r]);
    for (size_t r = 0; r < p->part.size() && r < M->subs.size(); ++r) free_single(M->subs[r], p->part[r]);
    delete p;
}

struct MultiGuard {
    pk_ctx *M;
    pk_dpoly *p;
    ~MultiGuard() { multi_free(M, p); }
    pk_dpoly *release()
    {
        pk_dpoly *r = p;
        p = nullptr;
        return r;
    }
};

pk_dpoly *multi_new(pk_ctx *M, u32 n_vars)
{
    pk_dpoly *p = new pk_dpoly;
    p->is_multi = true;
    p->n_vars = n_vars;
    p->rep.assign(M->subs.size(), nullptr);
    return p;
}

void multi_require(const pk_ctx *M, const pk_dpoly *p)
{
    if (!p) fail(PK_E_ARGS, "null polynomial");
    if (!p->is_multi || (p->rep.size() != M->subs.size())) fail(PK_E_ARGS, "polynomial does not belong to this multi-device context");
}

// a copy of `src` (complete, on device src_dev) on sub's device: peer-to-peer over NVLink, metadata from the host struct
pk_dpoly *replicate_single(pk_ctx *sub, const pk_dpoly *src, int src_dev)
{
    DPolyGuard d{sub, new_dpoly(sub, src->n, src->n_vars, src->limbs)};
    if (src->n) {
        CUDA_CHECK(cudaMemcpyPeerAsync(d.p->key, sub->device, src->key, src_dev, src->n * 8, sub->stream));
        for (u32 l = 0; l < src->limbs; ++l)
            CUDA_CHECK(cudaMemcpyPeerAsync(d.p->planes + (u64)l * d.p->stride, sub->device, src->planes + (u64)l * src->stride, src_dev,
                                           src->n * 8, sub->stream));
        CUDA_CHECK(cudaMemcpyPeerAsync(d.p->sign, sub->device, src->sign, src_dev, src->n, sub->stream));
    }
    d.p->has_meta = src->has_meta;
    d.p->emin = src->emin;
    d.p->emax = src->emax;
    d.p->egcd = src->egcd;
    d.p->max_bits = src->max_bits;
    d.p->any_neg = src->any_neg;
    d.p->any_zero = src->any_zero;
    return d.release();
}

// Host polynomial -> every device: ONE upload (device 0: host-to-device copy, exponent bounds, sort), then peer copies.
// (Measured on 8 GPUs: eight concurrent uploads of the same pageable host arrays contend in the driver, 0.8 ms against
// 0.35 ms for one.)
pk_dpoly *multi_upload(pk_ctx *M, const pk_poly *p)
{
    if (!p) fail(PK_E_ARGS, "null polynomial");
    MultiGuard g{M, multi_new(M, p->n_vars)};
    multi_run(M, [&](u32 r) {
        if (r == 0) g.p->rep[0] = upload_impl(M->subs[0], p);
    });
    multi_run(M, [&](u32 r) {
        if (r == 0) return;
        g.p->rep[r] = replicate_single(M->subs[r], g.p->rep[0], M->subs[0]->device);
        CUDA_CHECK(cudaStreamSynchronize(M->subs[r]->stream));
    });
    g.p->n = g.p->rep[0]->n;
    g.p->limbs = g.p->rep[0]->limbs;
    return g.release();
}

// Make the full polynomial resident on every device: the parts of a partitioned product are copied peer-to-peer into
// one array per device, in part order (= code order).
void multi_replicate(pk_ctx *M, pk_dpoly *p)
{
    bool have = true;
    for (pk_dpoly *q : p->rep) have = have && q;
    if (have) return;
    if (p->part.empty()) fail(PK_E_ARGS, "polynomial has neither replicas nor parts");
    u32 L = 1;
    for (pk_dpoly *q : p->part) L = std::max(L, q->limbs);
    const u64 n = p->n;
    multi_run(M, [&](u32 r) {
        if (p->rep[r]) return;
        pk_ctx *sub = M->subs[r];
        DPolyGuard d{sub, new_dpoly(sub, n, p->n_vars, L)};
        u64 off = 0;
        for (size_t q = 0; q < p->part.size(); ++q) {
            const pk_dpoly *src = p->part[q];
            const int sdev = M->subs[q]->device;
            if (!src->n) continue;
            CUDA_CHECK(cudaMemcpyPeerAsync(d.p->key + off, sub->device, src->key, sdev, src->n * 8, sub->stream));
            for (u32 l = 0; l < L; ++l) {
                u64 *dst = d.p->planes + (u64)l * d.p->stride + off;
                if (l < src->alloc_limbs)
                    CUDA_CHECK(cudaMemcpyPeerAsync(dst, sub->device, src->planes + (u64)l * src->stride, sdev, src->n * 8, sub->stream));
                else CUDA_CHECK(cudaMemsetAsync(dst, 0, src->n * 8, sub->stream));
            }
            CUDA_CHECK(cudaMemcpyPeerAsync(d.p->sign + off, sub->device, src->sign, sdev, src->n, sub->stream));
            off += src->n;
        }
        CUDA_CHECK(cudaStreamSynchronize(sub->stream));
        p->rep[r] = d.release();
    });
}

void multi_collect_stats(pk_ctx *M)
{
    pk_stats acc{};
    for (pk_ctx *sub : M->subs) {
        pk_stats s{};
        pk_get_stats(sub, &s);
        acc.n1 = s.n1;
        acc.n2 = s.n2;
        acc.term_products += s.term_products;
        acc.result_terms += s.result_terms;
        acc.code_range = s.code_range;
        acc.table_slots = s.table_slots;
        acc.algo = s.algo;
        acc.acc_lanes = s.acc_lanes;
        acc.chunk_bits = s.chunk_bits;
        acc.result_limbs = std::max(acc.result_limbs, s.result_limbs);
        acc.kernel_launches += s.kernel_launches;
        acc.retries = s.retries;
        acc.total_launches += s.total_launches;
        acc.mul_kernel_ms = std::max(acc.mul_kernel_ms, s.mul_kernel_ms); // the devices run side by side
        acc.device_ms = std::max(acc.device_ms, s.device_ms);
    }
    M->stats = acc;
}

pk_opts multi_opts(const pk_opts *opts_in)
{
    pk_opts o{};
    if (opts_in) {
        if (opts_in->struct_size != sizeof(pk_opts)) fail(PK_E_ARGS, "pk_opts.struct_size mismatch");
        o = *opts_in;
        if (o.part_count > 1) fail(PK_E_ARGS, "a multi-device context partitions the product itself (part_count must be 0 or 1)");
    }
    o.struct_size = sizeof(pk_opts);
    return o;
}

pk_dpoly *multi_mul_dev(pk_ctx *M, pk_dpoly *a, pk_dpoly *b, const pk_opts *opts_in)
{
    multi_require(M, a);
    multi_require(M, b);
    const pk_opts o0 = multi_opts(opts_in);
    multi_replicate(M, a);
    if (b != a) multi_replicate(M, b);
    const u32 N = (u32)M->subs.size();
    MultiGuard g{M, multi_new(M, a->n_vars)};
    g.p->rep.assign(N, nullptr);
    g.p->part.assign(N, nullptr);
    multi_run(M, [&](u32 r) {
        pk_opts o = o0;
        o.part_index = r;
        o.part_count = N;
        g.p->part[r] = mul_impl(M->subs[r], a->rep[r], b->rep[r], &o);
    });
    g.p->n = 0;
    g.p->limbs = 1;
    for (pk_dpoly *q : g.p->part) {
        g.p->n += q->n;
        if (q->n) g.p->limbs = std::max(g.p->limbs, q->limbs);
    }
    multi_collect_stats(M);
    return g.release();
}

// terms of `p` (on sub's device) -> host arrays with magnitude rows of L limbs; returns with the copies done
void copy_terms_to_host(pk_ctx *sub, const pk_dpoly *p, u32 L, int64_t *key, uint64_t *mag, int8_t *sign)
{
    const u64 n = p->n;
    if (!n) return;
    CUDA_CHECK(cudaMemcpyAsync(key, p->key, n * 8, cudaMemcpyDeviceToHost, sub->stream));
    if (L == 1) {
        CUDA_CHECK(cudaMemcpyAsync(mag, p->planes, n * 8, cudaMemcpyDeviceToHost, sub->stream));
        CUDA_CHECK(cudaMemcpyAsync(sign, p->sign, n, cudaMemcpyDeviceToHost, sub->stream));
        CUDA_CHECK(cudaStreamSynchronize(sub->stream));
    } else {
        DevBuf aos(sub, n * L * 8);
        PK_LAUNCH(sub, k_planes_to_aos, grid_for(n, 256), 256, 0, p->planes, aos.as<u64>(), n, std::min(p->alloc_limbs, L), p->stride, L);
        CUDA_CHECK(cudaMemcpyAsync(mag, aos.p, n * L * 8, cudaMemcpyDeviceToHost, sub->stream));
        CUDA_CHECK(cudaMemcpyAsync(sign, p->sign, n, cudaMemcpyDeviceToHost, sub->stream));
        CUDA_CHECK(cudaStreamSynchronize(sub->stream));
    }
}

struct HostLayout {
    PinnedBuf *pb = nullptr;
    int64_t *key = nullptr;
    uint64_t *mag = nullptr;
    int8_t *sign = nullptr;
};

HostLayout multi_host_result(pk_ctx *M, u64 n, u32 L, pk_result *out)
{
    HostLayout h;
    memset(out, 0, sizeof(*out));
    out->n = n;
    out->limbs = L;
    if (!n) return h;
    const size_t key_b = (n * 8 + 63) & ~63ull, mag_b = (n * L * 8 + 63) & ~63ull, sign_b = (n + 63) & ~63ull;
    DeviceScope ds(M->subs[0]->device);
    h.pb = pinned_acquire(M, key_b + mag_b + sign_b);
    char *base = static_cast<char *>(h.pb->ptr);
    h.key = out->key = reinterpret_cast<int64_t *>(base);
    h.mag = out->mag = reinterpret_cast<uint64_t *>(base + key_b);
    h.sign = out->sign = reinterpret_cast<int8_t *>(base + key_b + mag_b);
    out->owner = h.pb->ptr;
    return h;
}

void multi_download(pk_ctx *M, const pk_dpoly *p, pk_result *out)
{
    multi_require(M, p);
    const u32 N = (u32)M->subs.size();
    if (p->part.empty()) { // replicated: device 0 holds everything
        const pk_dpoly *src = p->rep[0];
        HostLayout h = multi_host_result(M, src->n, std::max<u32>(1, src->limbs), out);
        out->n_vars = p->n_vars;
        try {
            multi_run(M, [&](u32 r) {
                if (r == 0) copy_terms_to_host(M->subs[0], src, out->limbs, h.key, h.mag, h.sign);
            });
        } catch (...) {
            if (h.pb) h.pb->in_use = false;
            memset(out, 0, sizeof(*out));
            throw;
        }
        return;
    }
    std::vector<u64> off(N + 1, 0);
    for (u32 r = 0; r < N; ++r) off[r + 1] = off[r] + p->part[r]->n;
    const u32 L = std::max<u32>(1, p->limbs);
    HostLayout h = multi_host_result(M, off[N], L, out);
    out->n_vars = p->n_vars;
    if (!off[N]) return;
    try {
        // every device copies its part over its own PCIe link, straight to its offset of the one host result
        multi_run(M, [&](u32 r) {
            copy_terms_to_host(M->subs[r], p->part[r], L, h.key + off[r], h.mag + off[r] * L, h.sign + off[r]);
        });
    } catch (...) {
        if (h.pb) h.pb->in_use = false;
        memset(out, 0, sizeof(*out));
        throw;
    }
}

// host-to-host product on N devices: (upload, multiply part r) on every device at once, then the N downloads at once
void multi_mul(pk_ctx *M, const pk_poly *a, const pk_poly *b, const pk_opts *opts_in, pk_result *out)
{
    if (!a || !b) fail(PK_E_ARGS, "null operand");
    const pk_opts o0 = multi_opts(opts_in);
    const u32 N = (u32)M->subs.size();
    const bool trace = getenv("PK_MULTI_TRACE") != nullptr; // debug: wall time of the phases
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    std::vector<double> t_up(N, 0.0), t_mul(N, 0.0);
    MultiGuard g{M, multi_new(M, a->n_vars)};
    g.p->part.assign(N, nullptr);
    // the operands go to device 0 (and, when there are many devices, device 1 takes the second operand) and are
    // replicated peer-to-peer; then every device multiplies its part
    const u32 b_home = (N > 1 && a != b) ? 1u : 0u;
    struct HomeGuard { // (freed with the owning device current, whichever thread unwinds)
        pk_ctx *sub;
        pk_dpoly *p = nullptr;
        ~HomeGuard() { free_single(sub, p); }
    } a0{M->subs[0]}, b0{M->subs[b_home]};
    multi_run(M, [&](u32 r) {
        const auto w0 = std::chrono::steady_clock::now();
        if (r == 0) a0.p = upload_impl(M->subs[0], a);
        if (r == b_home && a != b) b0.p = upload_impl(M->subs[b_home], b);
        t_up[r] = ms_since(w0);
    });
    multi_run(M, [&](u32 r) {
        pk_ctx *sub = M->subs[r];
        const auto w0 = std::chrono::steady_clock::now();
        DPolyGuard da{sub, r == 0 ? nullptr : replicate_single(sub, a0.p, M->subs[0]->device)};
        DPolyGuard db{sub, (a == b || r == b_home) ? nullptr : replicate_single(sub, b0.p, M->subs[b_home]->device)};
        pk_dpoly *pa = r == 0 ? a0.p : da.p;
        pk_dpoly *pb = (a == b) ? pa : (r == b_home ? b0.p : db.p);
        pk_opts o = o0;
        o.part_index = r;
        o.part_count = N;
        g.p->part[r] = mul_impl(sub, pa, pb, &o);
        t_mul[r] = t_up[r] + ms_since(w0);
    });
    const double t_phase1 = ms_since(t0);
    g.p->n = 0;
    g.p->limbs = 1;
    for (pk_dpoly *q : g.p->part) {
        g.p->n += q->n;
        if (q->n) g.p->limbs = std::max(g.p->limbs, q->limbs);
    }
    multi_collect_stats(M);
    multi_download(M, g.p, out);
    if (trace) {
        fprintf(stderr, "[pk multi] upload+multiply %.3f ms, + download %.3f ms;  per device (upload, upload+multiply):", t_phase1, ms_since(t0));
        for (u32 r = 0; r < N; ++r) fprintf(stderr, " (%.3f, %.3f)", t_up[r], t_mul[r]);
        fprintf(stderr, "\n");
    }
}

u64 multi_digest(pk_ctx *M, const pk_dpoly *p)
{
    multi_require(M, p);
    const u32 N = (u32)M->subs.size();
    std::vector<uint64_t> d(N, 0);
    multi_run(M, [&](u32 r) {
        const pk_dpoly *q = p->part.empty() ? (r == 0 ? p->rep[0] : nullptr) : p->part[r];
        if (!q) return;
        if (pk_dpoly_digest(M->subs[r], q, &d[r]) != PK_OK) fail(PK_E_CUDA, M->subs[r]->err);
    });
    u64 s = 0;
    for (uint64_t x : d) s += x;
    return s;
}

void multi_destroy(pk_ctx *M)
{
    if (M->multi) {
        {
            std::lock_guard<std::mutex> lk(M->multi->m);
            M->multi->stop = true;
            M->multi->cv_job.notify_all();
        }
        for (auto &w : M->multi->w)
            if (w.th.joinable()) w.th.join();
        delete M->multi;
        M->multi = nullptr;
    }
    for (pk_ctx *sub : M->subs) pk_ctx_destroy(sub);
    M->subs.clear();
    for (auto &b : M->pinned)
        if (b.ptr) cudaFreeHost(b.ptr);
    delete M;
}

} // namespace
