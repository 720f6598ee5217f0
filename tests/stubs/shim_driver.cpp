// Test driver for the C++ drop-in shim.  Usage:
//   shim_driver probe                     -> checks the no-device contract, exit 0 if it holds
//   shim_driver run <in.bin> <out.bin>    -> runs ransacRegistration + icpRefine + GPURegistration::icpRefine
//   shim_driver pool <in.bin> <out.bin> <n_threads>  -> the same registration from n_threads worker threads at once (the
//          orchestrator's pool, src/pipeline.cpp:321-327); out.bin: per thread i32 device, then the 3 results of `run`,
//          then worldPose (16 f32) of the refined pose; last: u32 kept + kept x 16 f32 from filterDuplicates over all of them
// in.bin : u32 n_src, n_tgt, max_iter, icp_iter; f32 voxel, confidence, icp_thr; then src xyz, tgt xyz, tgt normals,
//          src desc, tgt desc (float32, packed).   out.bin: 3 x (16 f32 T column-major, fitness, rmse).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <thread>
#include <vector>
#ifdef B3D_REFERENCE_HEADERS
#include B3D_REFERENCE_HEADERS
#else
#include "registration.hpp"
#include "gpu_registration.hpp"
#endif
#include "b3d_registration_shim.hpp"
using namespace industry_picking;

static void put(FILE* f, const RegistrationResult& r) { fwrite(r.transformation.data(), 4, 16, f); fwrite(&r.fitness, 4, 1, f); fwrite(&r.rmse, 4, 1, f); }

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "probe")) {
        PointCloud a, b; a.points.assign(10, Eigen::Vector3f(0, 0, 0)); b.points = a.points;
        if (GPURegistration::isCudaAvailable()) { puts("cuda available"); return 0; }
        int thrown = 0;
        try { GPURegistration::icpRefine(a, b, Eigen::Matrix4f::Identity(), 0.01f); } catch (const std::runtime_error&) { thrown |= 1; }
        try { Registration::icpRefine(a, b, Eigen::Matrix4f::Identity(), 0.01f); } catch (const std::exception&) { thrown |= 2; }
        FPFHFeatures fa, fb; fa.descriptors.resize(10); fb.descriptors.resize(10);
        try { Registration::ransacRegistration(a, b, fa, fb, 0.001f); } catch (...) { thrown |= 4; }
        printf("no cuda: thrown mask %d\n", thrown);
        return thrown == 7 ? 0 : 1;            // no CPU fallback: every entry point must throw
    }
    if (argc == 4 && !strcmp(argv[1], "run")) {
        FILE* f = fopen(argv[2], "rb"); if (!f) return 2;
        uint32_t hdr[4]; float par[3];
        if (fread(hdr, 4, 4, f) != 4 || fread(par, 4, 3, f) != 3) return 2;
        PointCloud src, tgt; FPFHFeatures sf, tf;
        src.points.resize(hdr[0]); tgt.points.resize(hdr[1]); tgt.normals.resize(hdr[1]);
        sf.descriptors.resize(hdr[0]); tf.descriptors.resize(hdr[1]);
        bool ok = fread(src.points.data(), 12, hdr[0], f) == hdr[0] && fread(tgt.points.data(), 12, hdr[1], f) == hdr[1] &&
                  fread(tgt.normals.data(), 12, hdr[1], f) == hdr[1] && fread(sf.descriptors.data(), 132, hdr[0], f) == hdr[0] &&
                  fread(tf.descriptors.data(), 132, hdr[1], f) == hdr[1];
        fclose(f); if (!ok) return 2;
        try {
            RegistrationResult coarse = Registration::ransacRegistration(src, tgt, sf, tf, par[0], (int)hdr[2], par[1]);
            RegistrationResult fine = Registration::icpRefine(src, tgt, coarse.transformation, par[2], (int)hdr[3], true);
            RegistrationResult gfine = GPURegistration::icpRefine(src, tgt, coarse.transformation, par[2], (int)hdr[3]);
            FILE* o = fopen(argv[3], "wb"); if (!o) return 3;
            put(o, coarse); put(o, fine); put(o, gfine); fclose(o);
        } catch (const std::exception& e) { fprintf(stderr, "shim_driver: %s\n", e.what()); return 4; }
        return 0;
    }
    if (argc == 5 && !strcmp(argv[1], "pool")) {
        FILE* f = fopen(argv[2], "rb"); if (!f) return 2;
        uint32_t hdr[4]; float par[3];
        if (fread(hdr, 4, 4, f) != 4 || fread(par, 4, 3, f) != 3) return 2;
        PointCloud src, tgt; FPFHFeatures sf, tf;
        src.points.resize(hdr[0]); tgt.points.resize(hdr[1]); tgt.normals.resize(hdr[1]);
        sf.descriptors.resize(hdr[0]); tf.descriptors.resize(hdr[1]);
        bool ok = fread(src.points.data(), 12, hdr[0], f) == hdr[0] && fread(tgt.points.data(), 12, hdr[1], f) == hdr[1] &&
                  fread(tgt.normals.data(), 12, hdr[1], f) == hdr[1] && fread(sf.descriptors.data(), 132, hdr[0], f) == hdr[0] &&
                  fread(tf.descriptors.data(), 132, hdr[1], f) == hdr[1];
        fclose(f); if (!ok) return 2;
        const int nt = atoi(argv[4]);
        struct Out { int device = -1; RegistrationResult r[3]; Eigen::Matrix4f world; int err = 0; };
        std::vector<Out> outs(nt);
        Eigen::Matrix4f ext = Eigen::Matrix4f::Identity(); ext(0, 3) = 0.5f; ext(1, 3) = -0.25f; ext(2, 3) = 1.0f;
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t)
            pool.emplace_back([&, t] {
                try {
                    Out& o = outs[t];
                    o.r[0] = Registration::ransacRegistration(src, tgt, sf, tf, par[0], (int)hdr[2], par[1]);
                    o.r[1] = Registration::icpRefine(src, tgt, o.r[0].transformation, par[2], (int)hdr[3], true);
                    o.r[2] = GPURegistration::icpRefine(src, tgt, o.r[0].transformation, par[2], (int)hdr[3]);
                    o.world = b3d_shim::worldPose(ext, o.r[1].transformation);
                    o.device = b3d_shim::context_device();
                } catch (const std::exception& e) { fprintf(stderr, "shim_driver thread %d: %s\n", t, e.what()); outs[t].err = 1; }
            });
        for (auto& th : pool) th.join();
        std::vector<Eigen::Matrix4f> wps;
        for (const Out& o : outs) { if (o.err) return 4; wps.push_back(o.world); }
        std::vector<Eigen::Matrix4f> kept;
        try { kept = b3d_shim::filterDuplicates(wps, 0.01f); } catch (const std::exception& e) { fprintf(stderr, "shim_driver: %s\n", e.what()); return 4; }
        FILE* o = fopen(argv[3], "wb"); if (!o) return 3;
        for (const Out& r : outs) { fwrite(&r.device, 4, 1, o); put(o, r.r[0]); put(o, r.r[1]); put(o, r.r[2]); fwrite(r.world.data(), 4, 16, o); }
        uint32_t nk = (uint32_t)kept.size(); fwrite(&nk, 4, 1, o);
        for (const auto& k : kept) fwrite(k.data(), 4, 16, o);
        fclose(o);
        return 0;
    }
    return 64;
}
