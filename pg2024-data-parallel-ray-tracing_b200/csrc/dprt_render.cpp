// dprt_render.cpp -- the C++ host of the data-parallel path tracer: what Renderer::launch does in the reference
// (src/render/renderer.cpp:1576-2059: one MPI rank per GPU, scene tables, spp x runSample, image average +
// MPI_Reduce, image to disk), written against the C ABI of libdprt.so (include/dprt.h) and nothing else.
//
//   dprt_render --scene scene.dprt --out image.pfm [--spp 4] [--bounces 4] [--proxy 0|1] [--path-gen 0|1]
//               [--world W] [--devices D]   in-process rank group: W chunk owners driven by this thread, owner k on GPU k % D
//               [--nccl-id-file F]     one process per GPU: RANK / WORLD_SIZE / LOCAL_RANK from the environment (the
//                                      launcher of the reference is mpirun; here any launcher that sets them), rank 0
//                                      publishes the NCCL unique id through file F
//
// Scene file (written by pg2024-data-parallel-ray-tracing_b200/scene.py:save_scene; little endian):
//   "DPRTSCN1" | int32 nObjects nMaterials nLights | dprt_camera | dprt_material[nMaterials] | dprt_light_tri[nLights]
//   per object: dprt_object_desc (isProxy ignored) | int64 ntris | float verts[9 ntris] | float normals[9 ntris]
//               | int32 mats[ntris] | int64 visBytes | vis blob | int64 depthBytes | depth blob   (proxy MLP weights, may be 0)
// Output: PFM (RGB float32, bottom-up as the format demands) on the root rank, one JSON line of statistics on stdout.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "dprt.h"

namespace {

struct Object {
    dprt_object_desc desc;
    int64_t ntris = 0;
    std::vector<float> verts, normals;
    std::vector<int32_t> mats;
    std::vector<char> vis, depth;
};
struct Scene {
    dprt_camera cam;
    std::vector<dprt_material> materials;
    std::vector<dprt_light_tri> lights;
    std::vector<Object> objects;
};

bool read_exact(FILE* f, void* p, size_t n) { return n == 0 || fread(p, 1, n, f) == n; }

bool load_scene(const std::string& path, Scene& sc, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    char magic[8]; int32_t hdr[3];
    bool ok = read_exact(f, magic, 8) && std::memcmp(magic, "DPRTSCN1", 8) == 0 && read_exact(f, hdr, sizeof(hdr));
    if (ok) ok = hdr[0] >= 1 && hdr[0] <= 32 && hdr[1] >= 1 && hdr[1] <= DPRT_MAX_MATERIALS && hdr[2] >= 1 && hdr[2] <= DPRT_MAX_LIGHTS;
    if (ok) {
        sc.materials.resize(hdr[1]); sc.lights.resize(hdr[2]); sc.objects.resize(hdr[0]);
        ok = read_exact(f, &sc.cam, sizeof(sc.cam)) && read_exact(f, sc.materials.data(), sizeof(dprt_material) * hdr[1]) &&
             read_exact(f, sc.lights.data(), sizeof(dprt_light_tri) * hdr[2]);
    }
    for (size_t k = 0; ok && k < sc.objects.size(); k++) {
        Object& o = sc.objects[k];
        ok = read_exact(f, &o.desc, sizeof(o.desc)) && read_exact(f, &o.ntris, 8) && o.ntris >= 0 && o.ntris < (int64_t(1) << 27);
        if (!ok) break;
        o.verts.resize(9 * (size_t)o.ntris); o.normals.resize(9 * (size_t)o.ntris); o.mats.resize((size_t)o.ntris);
        int64_t vb = 0, db = 0;
        ok = read_exact(f, o.verts.data(), o.verts.size() * 4) && read_exact(f, o.normals.data(), o.normals.size() * 4) &&
             read_exact(f, o.mats.data(), o.mats.size() * 4) && read_exact(f, &vb, 8) && vb >= 0 && vb < (int64_t(1) << 30);
        if (ok) { o.vis.resize((size_t)vb); ok = read_exact(f, o.vis.data(), (size_t)vb) && read_exact(f, &db, 8) && db >= 0 && db < (int64_t(1) << 30); }
        if (ok) { o.depth.resize((size_t)db); ok = read_exact(f, o.depth.data(), (size_t)db); }
    }
    fclose(f);
    if (!ok) err = "malformed scene file " + path;
    return ok;
}

bool write_pfm(const std::string& path, const float* rgb, int w, int h) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "PF\n%d %d\n-1.0\n", w, h);
    for (int row = h - 1; row >= 0; row--) fwrite(rgb + (size_t)row * w * 3, sizeof(float), (size_t)w * 3, f);
    fclose(f);
    return true;
}

// the AccelerationStructure table of one rank (renderer.cpp:1812-1842): own objects as geometry, the others as proxies
int upload_scene(dprt_ctx* ctx, const Scene& sc, int rank) {
    int r;
    for (size_t k = 0; k < sc.objects.size(); k++) {
        const Object& o = sc.objects[k];
        dprt_object_desc d = o.desc;
        if (o.desc.nodeID == rank) {
            d.isProxy = 0;
            if ((r = dprt_upload_chunk(ctx, (int)k, &d, o.verts.data(), o.normals.data(), o.mats.data(), o.ntris))) return r;
        } else {
            d.isProxy = 1;
            if ((r = dprt_upload_proxy(ctx, (int)k, &d, o.vis.empty() ? nullptr : o.vis.data(), o.vis.size(),
                                       o.depth.empty() ? nullptr : o.depth.data(), o.depth.size()))) return r;
        }
    }
    if ((r = dprt_set_materials(ctx, sc.materials.data(), (int)sc.materials.size()))) return r;
    if ((r = dprt_set_lights(ctx, sc.lights.data(), (int)sc.lights.size()))) return r;
    return dprt_set_camera(ctx, &sc.cam);
}

int die(const char* what, dprt_ctx* ctx, int rc) {
    fprintf(stderr, "dprt_render: %s failed (%d): %s\n", what, rc, dprt_last_error(ctx));
    return 1;
}

// Samples in flight (dprt.h): `extra` contexts adopt `ctx`'s scene; context j renders samples j, j + K, ... from its own thread,
// then everything is accumulated into `ctx`. K = 1: the plain sample loop of Renderer::launch (renderer.cpp:1993-2022).
int render_samples(dprt_ctx* ctx, std::vector<dprt_ctx*>& extra, int spp, const char** what, dprt_ctx** failed) {
    const int K = 1 + (int)extra.size();
    std::vector<dprt_ctx*> all{ctx};
    all.insert(all.end(), extra.begin(), extra.end());
    std::vector<int> rc(K, 0);
    auto work = [&](int j) {
        for (int s = j; s < spp && !rc[j]; s += K) rc[j] = dprt_render_sample(all[j], s);
    };
    std::vector<std::thread> th;
    for (int j = 1; j < K; j++) th.emplace_back(work, j);
    work(0);
    for (auto& t : th) t.join();
    for (int j = 0; j < K; j++) if (rc[j]) { *what = "dprt_render_sample"; *failed = all[j]; return rc[j]; }
    for (int j = 1; j < K; j++) {
        const int r = dprt_accumulate_from(ctx, all[j]);
        if (r) { *what = "dprt_accumulate_from"; *failed = ctx; return r; }
    }
    return 0;
}

const char* arg(int argc, char** argv, const char* name, const char* dflt) {
    for (int i = 1; i + 1 < argc; i++) if (!std::strcmp(argv[i], name)) return argv[i + 1];
    return dflt;
}

}  // namespace

int main(int argc, char** argv) {
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);       // before the first CUDA call: one hardware queue per stream (--inflight)
    const std::string scenePath = arg(argc, argv, "--scene", ""), outPath = arg(argc, argv, "--out", "");
    if (scenePath.empty()) {
        fprintf(stderr, "usage: dprt_render --scene S.dprt [--out image.pfm] [--spp n] [--bounces n] [--proxy 0|1] [--path-gen 0|1] "
                        "[--inflight K] [--world W | --nccl-id-file F]\n");
        return 2;
    }
    Scene sc; std::string err;
    if (!load_scene(scenePath, sc, err)) { fprintf(stderr, "dprt_render: %s\n", err.c_str()); return 1; }
    const char* idFile = arg(argc, argv, "--nccl-id-file", nullptr);
    int world = std::atoi(arg(argc, argv, "--world", "0"));
    int rank = 0, local = 0;
    if (idFile) {
        rank = std::atoi(getenv("RANK") ? getenv("RANK") : "0");
        world = std::atoi(getenv("WORLD_SIZE") ? getenv("WORLD_SIZE") : "1");
        local = std::atoi(getenv("LOCAL_RANK") ? getenv("LOCAL_RANK") : "0");
    }
    int owners = 0;
    for (const Object& o : sc.objects) owners = o.desc.nodeID + 1 > owners ? o.desc.nodeID + 1 : owners;
    if (world <= 0) world = owners;
    if (world != owners) { fprintf(stderr, "dprt_render: scene has %d chunk owners, world is %d\n", owners, world); return 1; }

    dprt_config cfg; std::memset(&cfg, 0, sizeof(cfg));
    cfg.width = sc.cam.width; cfg.height = sc.cam.height;
    cfg.spp = std::atoi(arg(argc, argv, "--spp", "1"));
    cfg.bounces = std::atoi(arg(argc, argv, "--bounces", "4"));
    cfg.shadowPathCount = DPRT_DEFAULT_SHADOW_PATH_COUNT; cfg.maxCount = DPRT_DEFAULT_MAX_COUNT;   // renderer.cpp:1602-1603
    cfg.sceneSize = (int)sc.objects.size();
    cfg.proxyMode = std::atoi(arg(argc, argv, "--proxy", "0"));
    cfg.pathGenMode = std::atoi(arg(argc, argv, "--path-gen", world > 1 ? "1" : "0"));
    cfg.mlpDtype = std::atoi(arg(argc, argv, "--mlp-dtype", "1"));
    cfg.envColor[0] = 0.6f; cfg.envColor[1] = 0.7f; cfg.envColor[2] = 0.9f;
    const size_t N = (size_t)cfg.width * cfg.height;
    std::vector<float> image(3 * N);
    std::vector<dprt_ctx*> ctxs, extra;
    int inflight = std::atoi(arg(argc, argv, "--inflight", "1"));      // samples in flight per rank
    if (inflight < 1) inflight = 1;
    if (inflight > cfg.spp) inflight = cfg.spp;
    int r;
    // K - 1 more contexts of this rank on the same communicator, sharing the uploaded scene (collective across ranks)
    auto make_inflight = [&](dprt_ctx* ctx) -> int {
        if (world > 1 && !dprt_p2p_enabled(ctx)) return 0;          // the NCCL fallback exchange needs the communicator to itself
        for (int j = 1; j < inflight; j++) {
            dprt_ctx* c = nullptr;
            int rr = dprt_create_shared(&cfg, ctx, &c);
            if (rr) return rr;
            extra.push_back(c);
            if ((rr = dprt_adopt_scene(c, ctx))) return rr;
            if ((rr = dprt_reset_frame(c))) return rr;
        }
        return 0;
    };

    if (idFile && world > 1) {
        // ---- one process per GPU (the reference's deployment) ----
        char id[128];
        if (rank == 0) {
            if ((r = dprt_get_unique_id(id))) return die("dprt_get_unique_id", nullptr, r);
            const std::string tmp = std::string(idFile) + ".tmp";
            FILE* f = fopen(tmp.c_str(), "wb");
            if (!f || fwrite(id, 1, 128, f) != 128) { fprintf(stderr, "dprt_render: cannot write %s\n", tmp.c_str()); return 1; }
            fclose(f);
            std::rename(tmp.c_str(), idFile);
        } else {
            bool got = false;
            for (int tries = 0; tries < 6000 && !got; tries++) {
                FILE* f = fopen(idFile, "rb");
                if (f) { got = fread(id, 1, 128, f) == 128; fclose(f); }
                if (!got) std::this_thread::sleep_for(std::chrono::milliseconds(10));
            }
            if (!got) { fprintf(stderr, "dprt_render: rank %d never saw %s\n", rank, idFile); return 1; }
        }
        dprt_ctx* ctx = nullptr;
        if ((r = dprt_create(&cfg, rank, world, local, id, &ctx))) return die("dprt_create", nullptr, r);
        ctxs.push_back(ctx);
        if ((r = upload_scene(ctx, sc, rank))) return die("scene upload", ctx, r);
        if ((r = make_inflight(ctx))) return die("samples in flight", nullptr, r);
        const auto t0 = std::chrono::steady_clock::now();
        if ((r = dprt_reset_frame(ctx))) return die("dprt_reset_frame", ctx, r);
        { const char* what = ""; dprt_ctx* bad = ctx; if ((r = render_samples(ctx, extra, cfg.spp, &what, &bad))) return die(what, bad, r); }
        if ((r = dprt_reduce_image(ctx, 0, rank == 0 ? image.data() : nullptr))) return die("dprt_reduce_image", ctx, r);
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        dprt_stats st; dprt_get_stats(ctx, &st);
        printf("{\"rank\": %d, \"world\": %d, \"seconds\": %.6f, \"rays_walked\": %lld, \"paths_sent_offrank\": %lld, \"exchange_iters\": %lld}\n",
               rank, world, sec, (long long)st.rays_walked, (long long)st.paths_sent_offrank, (long long)st.exchange_iters);
    } else {
        // ---- in-process rank group (or a single rank) ----
        int ndev = std::atoi(arg(argc, argv, "--devices", "1"));      // chunk owner k runs on GPU k % ndev
        if (ndev < 1) ndev = 1;
        ctxs.resize(world, nullptr);
        for (int k = 0; k < world; k++) {
            if ((r = dprt_create(&cfg, k, world, k % ndev, nullptr, &ctxs[k]))) return die("dprt_create", nullptr, r);
            if ((r = upload_scene(ctxs[k], sc, k))) return die("scene upload", ctxs[k], r);
        }
        if (world == 1 && (r = make_inflight(ctxs[0]))) return die("samples in flight", nullptr, r);
        const auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < world; k++) if ((r = dprt_reset_frame(ctxs[k]))) return die("dprt_reset_frame", ctxs[k], r);
        if (world == 1) {
            const char* what = ""; dprt_ctx* bad = ctxs[0];
            if ((r = render_samples(ctxs[0], extra, cfg.spp, &what, &bad))) return die(what, bad, r);
        } else {
            for (int s = 0; s < cfg.spp; s++)
                if ((r = dprt_render_sample_group(ctxs.data(), world, s))) return die("dprt_render_sample_group", ctxs[0], r);
        }
        r = world == 1 ? dprt_reduce_image(ctxs[0], 0, image.data()) : dprt_reduce_image_group(ctxs.data(), world, 0, image.data());
        if (r) return die("dprt_reduce_image", ctxs[0], r);
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        long long walked = 0, sent = 0;
        for (int k = 0; k < world; k++) { dprt_stats st; dprt_get_stats(ctxs[k], &st); walked += st.rays_walked; sent += st.paths_sent_offrank; }
        printf("{\"rank\": 0, \"world\": %d, \"seconds\": %.6f, \"rays_walked\": %lld, \"paths_sent_offrank\": %lld}\n", world, sec, walked, sent);
    }
    if (rank == 0 && !outPath.empty() && !write_pfm(outPath, image.data(), cfg.width, cfg.height)) {
        fprintf(stderr, "dprt_render: cannot write %s\n", outPath.c_str());
        return 1;
    }
    for (dprt_ctx* c : extra) dprt_destroy(c);         // borrowers before the owner of the scene and the communicator
    for (dprt_ctx* c : ctxs) dprt_destroy(c);
    return 0;
}
