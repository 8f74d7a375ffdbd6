// Several GPUs behind one handle (SURVEY.md §8b "one context per process owning G devices"; §8e): the counterpart of the
// reference's RenderManager driving all of its workers from one shared tile queue (renderer/render_manager.rs:78-97,
// 197-236; render_worker.rs:172-198 `pop_tile_or_signal_finish`).
//   * one yk_context + one host worker thread per device, the scene replicated: the first device validates, uploads and
//     repacks the caller's one host copy, the others pull the repacked scene from it over NVLink (scene_clone_impl);
//   * the tile list is consumed through ONE shared cursor by guided self-scheduling: a worker pops the next run of tiles (a
//     range of a strided order of the spiral list, so that each run covers the film centre-out and costs about the same per
//     tile; remaining / 2G tiles, i.e. large runs first and small ones last), renders it with the wavefront pipeline of its
//     device and comes back for more — a slower or busier device simply takes fewer tiles, so no device idles while another
//     still has a backlog;
//   * no gather step: the film lives on the first device and every other device's film kernels (k_film_store / k_film_add,
//     and the primary-hit id image) store their finished pixels straight into it through peer mappings, i.e. as NVLink
//     writes overlapped with the rendering of the following batches. Without peer access the devices render into local
//     films whose rendered pixels are copied over to the first device at the end.
// Accumulating films add a pixel once per sample index; to keep the order of those adds fixed (the reference's is
// whatever its workers' timing makes it) a tile always goes to device `tile.index mod G` there.
// Part of the translation unit render.cu (uses render_impl / scene_create_impl).

struct yk_multi {
    std::vector<yk_context*> ctx;     // ctx[0] owns the film
    std::vector<int> peer_ok;         // device i can store into device 0's memory
    std::mutex mu;                    // one render at a time
    float* merge_stage = nullptr;     // device 0: staging for the fallback gather (+ the initial film of accumulating renders)
    size_t merge_cap = 0;
};
struct yk_multi_scene {
    yk_multi* owner = nullptr;
    std::vector<yk_scene*> scene;
};

namespace {

// Gather of a device that could not store into the first device's film: the pixels the device wrote replace the film's, every
// other pixel is left alone (like Film::update_tile touches only its tile, film.rs:260-279). Averaging films: the local film
// started as kUnrendered (a NaN payload no kernel produces). Accumulating films: the local film started as a copy of the
// initial film, so that a pixel's adds associate exactly as on one device ((film + s0) + s1 ...); a pixel is the device's own
// iff it differs from that initial film (an unchanged own pixel needs no copy).
constexpr uint32_t kUnrendered = 0x7fc0dead;
__global__ void k_film_merge(float* film, const float* part, const float* initial, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = __float_as_uint(part[i]);
    if (v != (initial ? __float_as_uint(initial[i]) : kUnrendered)) film[i] = part[i];
}
__global__ void k_fill_u32(uint32_t* p, size_t n, uint32_t v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_ids_merge(int32_t* ids, const int32_t* part, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && part[i] >= 0) ids[i] = part[i];
}

void sum_stats(yk_stats* a, const yk_stats& b) {
    a->ray_count += b.ray_count; a->shadow_rays += b.shadow_rays; a->samples += b.samples;
    a->closest_nodes += b.closest_nodes; a->closest_tris += b.closest_tris; a->any_nodes += b.any_nodes; a->any_tris += b.any_tris;
    a->primary_hit_hash += b.primary_hit_hash;
    a->device_ms += b.device_ms; a->trace_closest_ms += b.trace_closest_ms; a->trace_any_ms += b.trace_any_ms; a->shade_ms += b.shade_ms;
    a->kernel_launches += b.kernel_launches; a->trace_closest_launches += b.trace_closest_launches;
}

}  // namespace

extern "C" {

int yk_multi_create(const int* device_ids, int n_devices, yk_multi** out) {
    return yk_guard("yk_multi_create", [&]() -> int {
    if (!device_ids || n_devices < 1 || !out) return yk_set_error(YK_ERR_INVALID, "yk_multi_create: null argument or no device");
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (device_ids[i] == device_ids[j]) return yk_set_error(YK_ERR_INVALID, "yk_multi_create: a device is listed twice");
    auto m = std::make_unique<yk_multi>();
    auto fail = [&](int rc) {
        for (yk_context* c : m->ctx) yk_context_destroy(c);
        return rc;
    };
    for (int i = 0; i < n_devices; ++i) {
        yk_context* c = nullptr;
        const int rc = yk_context_create(device_ids[i], &c);
        if (rc != YK_OK) return fail(rc);
        m->ctx.push_back(c);
    }
    m->peer_ok.assign((size_t)n_devices, 0);
    m->peer_ok[0] = 1;
    for (int i = 1; i < n_devices; ++i) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, device_ids[i], device_ids[0]) == cudaSuccess && can) {
            if (cudaSetDevice(device_ids[i]) != cudaSuccess) continue;
            const cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[0], 0);
            if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) m->peer_ok[(size_t)i] = 1;
            (void)cudaGetLastError();
            // ... and the other way round, so that the scene clone's peer copies (device 0 -> i) run device to device on every
            // driver instead of being staged through the host
            int back = 0;
            if (cudaDeviceCanAccessPeer(&back, device_ids[0], device_ids[i]) == cudaSuccess && back && cudaSetDevice(device_ids[0]) == cudaSuccess)
                (void)cudaDeviceEnablePeerAccess(device_ids[i], 0);
            (void)cudaGetLastError();
            // Scenes live in device 0's stream-ordered memory pool (dev_alloc), whose memory cudaDeviceEnablePeerAccess does not
            // map: without this grant the clone's 1 GB of peer copies were staged through the host at 23 GB/s.
            if (m->peer_ok[(size_t)i]) {
                cudaMemPool_t pool = nullptr;
                if (cudaDeviceGetDefaultMemPool(&pool, device_ids[0]) == cudaSuccess) {
                    cudaMemAccessDesc desc{};
                    desc.location.type = cudaMemLocationTypeDevice;
                    desc.location.id = device_ids[i];
                    desc.flags = cudaMemAccessFlagsProtReadWrite;
                    (void)cudaMemPoolSetAccess(pool, &desc, 1);
                }
                (void)cudaGetLastError();
            }
        }
    }
    if (getenv("YK_MULTI_NO_PEER"))  // development / tests: force the gather fallback
        for (int i = 1; i < n_devices; ++i) m->peer_ok[(size_t)i] = 0;
    *out = m.release();
    return YK_OK;
    });
}

void yk_multi_destroy(yk_multi* m) {
    if (!m) return;
    if (m->merge_stage) {
        cudaSetDevice(m->ctx[0]->device);
        cudaFree(m->merge_stage);
    }
    for (yk_context* c : m->ctx) yk_context_destroy(c);
    delete m;
}

int yk_multi_device_count(const yk_multi* m) { return m ? (int)m->ctx.size() : 0; }
yk_context* yk_multi_context(yk_multi* m, int i) { return (m && i >= 0 && (size_t)i < m->ctx.size()) ? m->ctx[(size_t)i] : nullptr; }
int yk_multi_peer_stores(const yk_multi* m, int i) { return (m && i >= 0 && (size_t)i < m->ctx.size()) ? m->peer_ok[(size_t)i] : 0; }

int yk_multi_scene_create(yk_multi* m, const yk_scene_desc* d, yk_multi_scene** out) {
    return yk_guard("yk_multi_scene_create", [&]() -> int {
    if (!m || !d || !out) return yk_set_error(YK_ERR_INVALID, "yk_multi_scene_create: null argument");
    auto ms = std::make_unique<yk_multi_scene>();
    ms->owner = m;
    ms->scene.assign(m->ctx.size(), nullptr);
    // device 0 validates (host threads, overlapped with its upload); the others reuse the verdict
    SceneCheck check;
    int rc = scene_create_impl(m->ctx[0], d, nullptr, &check, &ms->scene[0]);
    if (rc != YK_OK) return rc;
    const bool no_clone = getenv("YK_MULTI_NO_CLONE") != nullptr;  // development / tests: every device uploads for itself
    std::vector<int> rcs(m->ctx.size(), YK_OK);
    std::vector<std::string> errs(m->ctx.size());
    std::vector<std::thread> th;
    for (size_t i = 1; i < m->ctx.size(); ++i)
        th.emplace_back([&, i] {
            rcs[i] = yk_guard("yk_multi_scene_create", [&]() -> int {
                // a device with a peer mapping of device 0 pulls the repacked scene from there (NVLink) instead of uploading and
                // repacking the host arrays again
                if (m->peer_ok[i] && !no_clone) return scene_clone_impl(m->ctx[i], ms->scene[0], &ms->scene[i]);
                return scene_create_impl(m->ctx[i], d, &check, nullptr, &ms->scene[i]);
            });
            if (rcs[i] != YK_OK) errs[i] = yk_last_error();  // (the message is thread-local)
        });
    for (auto& t : th) t.join();
    for (size_t i = 1; i < m->ctx.size(); ++i)
        if (rcs[i] != YK_OK) {
            for (yk_scene* s : ms->scene) yk_scene_destroy(s);
            return yk_set_error(rcs[i], errs[i]);
        }
    *out = ms.release();
    return YK_OK;
    });
}

void yk_multi_scene_destroy(yk_multi_scene* ms) {
    if (!ms) return;
    for (yk_scene* s : ms->scene) yk_scene_destroy(s);
    delete ms;
}

int yk_multi_render(yk_multi* m, const yk_multi_scene* ms, const yk_camera* cam, const yk_film_settings* fs, const yk_sampler* sm,
                    const yk_integrator* in, const yk_tile* tiles, uint32_t n_tiles, const yk_render_opts* opts, float* film_rgb,
                    yk_stats* stats, yk_stats* per_device) {
    return yk_guard("yk_multi_render", [&]() -> int {
    const auto wall0 = std::chrono::steady_clock::now();
    if (!m || !ms || !cam || !fs || !sm || !in || !film_rgb) return yk_set_error(YK_ERR_INVALID, "yk_multi_render: null argument");
    if (ms->owner != m) return yk_set_error(YK_ERR_INVALID, "yk_multi_render: scene belongs to another device group");
    if (n_tiles && !tiles) return yk_set_error(YK_ERR_INVALID, "yk_multi_render: null tile list");
    if (!fs->res_x || !fs->res_y || fs->res_x > 0xffffu || fs->res_y > 0xffffu)
        return yk_set_error(YK_ERR_INVALID, "yk_multi_render: film resolution must fit u16 (integrators/mod.rs:140-141)");
    const size_t G = m->ctx.size();
    if (G == 1 && !getenv("YK_MULTI_FORCE_CURSOR")) {  // nothing to share (tests force the cursor path on one-GPU boxes)
        const int rc = render_impl(m->ctx[0], ms->scene[0], cam, fs, sm, in, tiles, n_tiles, opts, film_rgb, stats, nullptr, nullptr);
        if (rc == YK_OK && stats && per_device) per_device[0] = *stats;
        return rc;
    }
    std::lock_guard<std::mutex> guard(m->mu);
    yk_context* root = m->ctx[0];
    std::lock_guard<std::recursive_mutex> root_guard(root->mu);
    CUDA_TRY(cudaSetDevice(root->device));
    const uint32_t flags = opts ? opts->flags : 0u;
    const bool on_device = (flags & YK_RENDER_FILM_ON_DEVICE) != 0;
    const bool accumulate = fs->accumulate != 0;
    const size_t n_pixels = (size_t)fs->res_x * fs->res_y, film_bytes = n_pixels * 3 * sizeof(float);
    const uint32_t spp = sm->kind == YK_SAMPLER_UNIFORM ? sm->nx : sm->nx * sm->ny;

    // The film (and the optional hit-id image) on device 0.
    if (root->film_cap < n_pixels) {
        cudaFree(root->d_accum); cudaFree(root->d_film); cudaFree(root->d_hit_ids);
        root->d_accum = nullptr; root->d_film = nullptr; root->d_hit_ids = nullptr;
        root->film_cap = 0;
        CUDA_TRY(cudaMalloc((void**)&root->d_accum, film_bytes));
        CUDA_TRY(cudaMalloc((void**)&root->d_film, film_bytes));
        CUDA_TRY(cudaMalloc((void**)&root->d_hit_ids, n_pixels * sizeof(int32_t)));
        root->film_cap = n_pixels;
    }
    float* d_film = on_device ? film_rgb : root->d_film;
    unsigned long long area = 0;
    for (uint32_t t = 0; t < n_tiles; ++t) {
        const yk_tile& tl = tiles[t];
        if (tl.x0 >= tl.x1 || tl.y0 >= tl.y1 || tl.x1 > fs->res_x || tl.y1 > fs->res_y)
            return yk_set_error(YK_ERR_INVALID, "yk_multi_render: tile outside the film (film.rs:224-231)");
        area += (unsigned long long)(tl.x1 - tl.x0) * (tl.y1 - tl.y0);
    }
    cudaStream_t s0 = root->stream;
    if (!on_device) {
        if (!accumulate && area == n_pixels) CUDA_TRY(cudaMemsetAsync(d_film, 0, film_bytes, s0));
        else CUDA_TRY(cudaMemcpyAsync(d_film, film_rgb, film_bytes, cudaMemcpyHostToDevice, s0));
    }
    int32_t* d_ids = nullptr;
    if (opts && opts->hit_ids) {
        d_ids = on_device ? opts->hit_ids : root->d_hit_ids;
        k_fill_i32<<<(unsigned)((n_pixels + 255) / 256), 256, 0, s0>>>(d_ids, n_pixels, -1);
    }
    CUDA_TRY(cudaStreamSynchronize(s0));

    bool any_gather = false;
    for (size_t i = 1; i < G; ++i) any_gather = any_gather || !m->peer_ok[i];
    if (any_gather) {
        if (m->merge_cap < 2 * film_bytes) {
            cudaFree(m->merge_stage);
            m->merge_stage = nullptr;
            m->merge_cap = 0;
            CUDA_TRY(cudaMalloc((void**)&m->merge_stage, 2 * film_bytes));
            m->merge_cap = 2 * film_bytes;
        }
        if (accumulate) CUDA_TRY(cudaMemcpy(m->merge_stage, d_film, film_bytes, cudaMemcpyDeviceToDevice));  // the initial film
    }
    // Devices without a peer mapping render into a local film that is gathered at the end.
    std::vector<float*> local_film(G, nullptr);
    std::vector<int32_t*> local_ids(G, nullptr);
    for (size_t i = 1; i < G; ++i) {
        if (m->peer_ok[i]) continue;
        yk_context* c = m->ctx[i];
        std::lock_guard<std::recursive_mutex> g2(c->mu);
        CUDA_TRY(cudaSetDevice(c->device));
        if (c->film_cap < n_pixels) {
            cudaFree(c->d_accum); cudaFree(c->d_film); cudaFree(c->d_hit_ids);
            c->d_accum = nullptr; c->d_film = nullptr; c->d_hit_ids = nullptr;
            c->film_cap = 0;
            CUDA_TRY(cudaMalloc((void**)&c->d_accum, film_bytes));
            CUDA_TRY(cudaMalloc((void**)&c->d_film, film_bytes));
            CUDA_TRY(cudaMalloc((void**)&c->d_hit_ids, n_pixels * sizeof(int32_t)));
            c->film_cap = n_pixels;
        }
        if (accumulate) CUDA_TRY(cudaMemcpyPeer(c->d_film, c->device, d_film, root->device, film_bytes));
        else k_fill_u32<<<(unsigned)((n_pixels * 3 + 255) / 256), 256, 0, c->stream>>>((uint32_t*)c->d_film, n_pixels * 3, kUnrendered);
        local_film[i] = c->d_film;
        if (d_ids) {
            k_fill_i32<<<(unsigned)((n_pixels + 255) / 256), 256, 0, c->stream>>>(c->d_hit_ids, n_pixels, -1);
            local_ids[i] = c->d_hit_ids;
        }
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }

    // Work distribution. Non-accumulating: runs of tiles popped from one cursor, *guided self-scheduling*: a pop takes
    // remaining / 2G tiles, so the first runs are large (half a device's fair share: few synchronisations and un-overlapped
    // kernel tails — measured on 2 GPUs, runs of a quarter batch made the 16-spp large scene 35 % slower than one call per
    // device) and the last ones small (the finish times differ by at most one small run, whatever the devices' speeds),
    // never below ~16 Mi paths (two wavefront batches of 8 Mi, one per pipe; with 64 Mi the two devices of a 0.4 s render
    // finished 39 ms apart). A job whose fair share is below 256 Mi paths is split statically, one run per device. The runs
    // are ranges of a *strided* order of the tile list (position j -> tile j * stride mod n_tiles, stride ~ 0.618 n_tiles and
    // coprime to it), so every run samples the whole spiral and costs about the same per tile. Accumulating: tile.index mod G
    // (fixed add order per pixel).
    uint32_t run_floor = 1, run_fixed = 0, stride = 1;
    if (!accumulate && n_tiles) {
        const unsigned long long paths_per_tile = std::max<unsigned long long>(1, area / n_tiles) * spp;
        run_floor = (uint32_t)std::max<unsigned long long>(1, (16ull << 20) / paths_per_tile);
        const uint32_t fair = (uint32_t)((n_tiles + G - 1) / G);
        // Short jobs are split statically: the devices of a box are alike, and the last guided runs (one floor = ~13 ms of a
        // B200) would leave them up to that far apart — 12 % of an 8-GPU render of 0.1 s, 0.3 % of the 4.6 s of the 4K config.
        unsigned long long static_below = 256ull << 20;  // paths per device
        if (const char* e = getenv("YK_MULTI_STATIC_BELOW")) static_below = strtoull(e, nullptr, 10);
        if (fair < 2 * run_floor || (unsigned long long)fair * paths_per_tile < static_below) run_fixed = fair;
        if (const char* e = getenv("YK_MULTI_RUN_TILES")) run_fixed = (uint32_t)std::max(1, atoi(e));
        auto gcd = [](uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; };
        stride = std::max(1u, (uint32_t)(0.6180339887 * n_tiles));
        while (stride > 1 && gcd(stride, n_tiles) != 1) --stride;
    }
    std::vector<std::vector<yk_tile>> fixed(accumulate ? G : 0);
    if (accumulate)
        for (uint32_t t = 0; t < n_tiles; ++t) fixed[tiles[t].index % G].push_back(tiles[t]);

    std::atomic<uint32_t> cursor{0};
    std::atomic<int> first_error{YK_OK};
    std::atomic<bool> stop{false};
    std::mutex progress_mu;
    std::string error_text;
    uint64_t done_samples = 0;
    const uint64_t total_samples = (uint64_t)area * (accumulate ? 1u : spp);
    std::vector<yk_stats> dev_stats(G);
    for (auto& st : dev_stats) st = yk_stats{};

    auto worker = [&](size_t i) {
        yk_context* c = m->ctx[i];
        yk_render_opts o{};
        if (opts) o = *opts;
        o.flags = (o.flags | YK_RENDER_FILM_ON_DEVICE | kRenderAuxInitialised);
        o.progress = nullptr;
        o.progress_user = nullptr;
        float* film_i = (i == 0 || m->peer_ok[i]) ? d_film : local_film[i];
        o.hit_ids = d_ids ? ((i == 0 || m->peer_ok[i]) ? d_ids : local_ids[i]) : nullptr;
        auto render_run = [&](const yk_tile* t, uint32_t n) {
            yk_stats st{};
            const int rc = yk_guard("yk_multi_render", [&]() -> int {
                return render_impl(c, ms->scene[i], cam, fs, sm, in, t, n, &o, film_i, &st, nullptr, nullptr);
            });
            if (rc != YK_OK) {
                int expected = YK_OK;
                std::lock_guard<std::mutex> g(progress_mu);
                if (first_error.compare_exchange_strong(expected, rc)) error_text = yk_last_error();
                stop = true;
                return;
            }
            sum_stats(&dev_stats[i], st);
            if (opts && opts->progress) {
                std::lock_guard<std::mutex> g(progress_mu);
                done_samples += st.samples;
                if (opts->progress(opts->progress_user, done_samples, total_samples)) stop = true;
            }
        };
        if (accumulate) {
            if (!fixed[i].empty()) render_run(fixed[i].data(), (uint32_t)fixed[i].size());
            return;
        }
        std::vector<yk_tile> mine;
        std::vector<uint32_t> idx;
        while (!stop) {
            uint32_t lo = cursor.load(), n_take = 0;
            do {  // guided pop
                if (lo >= n_tiles) break;
                const uint32_t remaining = n_tiles - lo;
                n_take = run_fixed ? run_fixed : std::max(run_floor, (uint32_t)((remaining + 2 * G - 1) / (2 * G)));
                n_take = std::min(n_take, remaining);
            } while (!cursor.compare_exchange_weak(lo, lo + n_take));
            if (lo >= n_tiles) break;
            idx.clear();
            for (uint32_t j = lo; j < lo + n_take; ++j) idx.push_back((uint32_t)(((uint64_t)j * stride) % n_tiles));
            std::sort(idx.begin(), idx.end());  // list (spiral) order inside the run
            mine.clear();
            for (uint32_t t : idx) mine.push_back(tiles[t]);
            render_run(mine.data(), (uint32_t)mine.size());
        }
    };
    {
        std::vector<std::thread> th;
        for (size_t i = 1; i < G; ++i) th.emplace_back(worker, i);
        worker(0);
        for (auto& t : th) t.join();
    }
    if (first_error != YK_OK) return yk_set_error(first_error, error_text);
    if (stop) return yk_set_error(YK_ERR_CANCELLED, "yk_multi_render: cancelled by the progress callback");

    // Gather of the devices that could not store into device 0's film.
    CUDA_TRY(cudaSetDevice(root->device));
    for (size_t i = 1; i < G; ++i) {
        if (m->peer_ok[i]) continue;
        float* stage = m->merge_stage + n_pixels * 3;
        CUDA_TRY(cudaMemcpyPeerAsync(stage, root->device, local_film[i], m->ctx[i]->device, film_bytes, s0));
        k_film_merge<<<(unsigned)((n_pixels * 3 + 255) / 256), 256, 0, s0>>>(d_film, stage, accumulate ? m->merge_stage : nullptr, n_pixels * 3);
        if (d_ids) {
            CUDA_TRY(cudaMemcpyPeerAsync(stage, root->device, local_ids[i], m->ctx[i]->device, n_pixels * sizeof(int32_t), s0));
            k_ids_merge<<<(unsigned)((n_pixels + 255) / 256), 256, 0, s0>>>(d_ids, (const int32_t*)stage, n_pixels);
        }
    }
    if (!on_device) {
        CUDA_TRY(cudaMemcpyAsync(film_rgb, d_film, film_bytes, cudaMemcpyDeviceToHost, s0));
        if (d_ids) CUDA_TRY(cudaMemcpyAsync(opts->hit_ids, d_ids, n_pixels * sizeof(int32_t), cudaMemcpyDeviceToHost, s0));
    }
    CUDA_TRY(cudaStreamSynchronize(s0));
    CUDA_TRY(cudaGetLastError());
    if (per_device)
        for (size_t i = 0; i < G; ++i) per_device[i] = dev_stats[i];
    if (stats) {
        yk_stats total{};
        double busiest = 0;
        for (const yk_stats& st : dev_stats) {
            sum_stats(&total, st);
            busiest = std::max(busiest, st.device_ms);
        }
        total.device_ms = busiest;  // the devices run side by side: the render took as long as the busiest one
        total.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
        *stats = total;
    }
    return YK_OK;
    });
}

}  // extern "C"
