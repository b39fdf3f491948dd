// tracer_cli.cpp — the reference's host flow (src/main.cpp:417-695) with the GPU renderer in the
// place of its accelerator call:  arg parse -> model::loadobj -> camera -> render -> PPM.
// Same flags as the reference where they exist (-m -o -v -l, main.cpp:467-529); `-w W,H` really sets
// the window here (in the reference it overwrites `look`, main.cpp:515-529); `--cuda` is accepted
// for symmetry with `--ispc`.  The CPU modes (--thread, --bvh, --ispc) do not exist in this build.
// `--gpus N` renders the frame on N GPUs of the box (interleaved 8-row bands, NCCL gather: tracer_cuda_init_multi);
// `--spp S` = S stratified jittered samples per pixel (extension); `--sorted` runs the reference's optional
// flatten+sort pass (src/simplify/flatten.cpp:50-82) before rendering.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/tracer_host.h"

static void parse_triple(const char *s, float out[3]) {
    if (std::sscanf(s, "%f,%f,%f", &out[0], &out[1], &out[2]) != 3) throw std::runtime_error(std::string("expected x,y,z: ") + s);
}

int main(int argc, char *argv[]) {
    try {
        std::string model, output;
        float eye[3] = {0, 1, 3}, look[3] = {0, 1, 0}; // main.cpp:426
        int W = 1024, H = 768;                          // main.cpp:427
        int device = 0, p6 = 0, bundle_cull = 0, gpus = 1, spp = 0;
        bool sorted = false;
        bool have_seed = false;
        unsigned seed = 0;
        int rng = TRACER_RNG_MT19937; // the serial path's generator
        for (int a = 1; a < argc; ++a) {
            const std::string f = argv[a];
            auto need = [&](const char *what) -> const char * {
                if (a + 1 >= argc) throw std::runtime_error(std::string("missing value for ") + what);
                return argv[++a];
            };
            if (f == "--cuda" || f == "--debug" || f == "--trace" || f == "--test") continue;
            if (f == "--thread" || f == "--bvh" || f == "--ispc")
                throw std::runtime_error(f + ": CPU/ISPC modes are not part of this build (GPU renderer only)");
            if (f == "-m") model = need("-m");
            else if (f == "-o") output = need("-o");
            else if (f == "-v") parse_triple(need("-v"), eye);
            else if (f == "-l") parse_triple(need("-l"), look);
            else if (f == "-w") {
                if (std::sscanf(need("-w"), "%d,%d", &W, &H) != 2) throw std::runtime_error("expected -w W,H");
            } else if (f == "--seed") seed = (unsigned)std::strtoul(need("--seed"), nullptr, 10), have_seed = true;
            else if (f == "--rng") {
                const std::string r = need("--rng");
                rng = r == "hash" ? TRACER_RNG_HASH : TRACER_RNG_MT19937;
            } else if (f == "--device") device = std::atoi(need("--device"));
            else if (f == "--p6") p6 = 1;
            else if (f == "--cull") bundle_cull = 3; // optional bundle-cull mode: same bytes out, much faster on big scenes
            else if (f == "--gpus") gpus = std::atoi(need("--gpus"));
            else if (f == "--spp") spp = std::atoi(need("--spp"));
            else if (f == "--sorted") sorted = true;
            else throw std::runtime_error("Unknown argument: " + f); // main.cpp:531-534
        }
        tracer_scene_host *scene = nullptr;
        tracer_scene_flat empty{};
        const int32_t zero_off[1] = {0};
        empty.geom_tri_offset = zero_off;
        const tracer_scene_flat *flat = &empty; // no -m: empty scene, black image (main.cpp:537-543)
        if (!model.empty()) {
            if (tracer_scene_load_obj(model.c_str(), &scene)) throw std::runtime_error(tracer_host_last_error());
            flat = tracer_scene_host_flat(scene);
        }
        tracer_scene_host *sorted_scene = nullptr;
        if (sorted && scene) { // main.cpp:566-567 runs flatten_scene before rendering
            if (tracer_scene_flatten_sorted(flat, &sorted_scene)) throw std::runtime_error(tracer_host_last_error());
            flat = tracer_scene_host_flat(sorted_scene);
        }
        tracer_camera cam;
        const float vup[3] = {0, 1, 0};
        tracer_camera_lookat(eye, look, vup, 60.f, float(W) / H, &cam); // main.cpp:548-551
        if (gpus > 1) {
            if (rng == TRACER_RNG_MT19937) {
                if (have_seed) throw std::runtime_error("--gpus > 1 needs --rng hash: the mt19937 stream is sequential over the whole frame (main.cpp:587-589)");
                rng = TRACER_RNG_HASH; // unseeded runs are random anyway (main.cpp:587-588)
            }
            if (tracer_cuda_init_multi(gpus)) throw std::runtime_error(tracer_cuda_last_error());
        } else if (tracer_cuda_init(device)) throw std::runtime_error(tracer_cuda_last_error());
        tracer_render_opts o{};
        o.struct_size = sizeof o;
        o.rng_mode = rng;
        o.bundle_cull = bundle_cull;
        o.samples_per_pixel = spp;
        if (spp > 1 && rng == TRACER_RNG_MT19937) o.rng_mode = rng = TRACER_RNG_HASH; // the serial replay is 1 spp only
        o.seed = have_seed ? seed : std::random_device{}(); // main.cpp:587-588
        std::vector<uint8_t> rgb((size_t)W * H * 3);
        const auto t0 = std::chrono::high_resolution_clock::now(); // main.cpp:583
        if (tracer_cuda_render(flat, &cam, W, H, &o, rgb.data())) throw std::runtime_error(tracer_cuda_last_error());
        const auto t1 = std::chrono::high_resolution_clock::now();
        std::cerr << "\n CUDA      : true" << std::endl;
        std::cerr << "\n GPUs      : " << gpus << std::endl;
        std::cerr << "\n Duration  : " << std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count() << std::endl;
        if (!output.empty()) {
            if (tracer_write_ppm(output.c_str(), rgb.data(), W, H, p6)) throw std::runtime_error(tracer_host_last_error());
            std::cout << "Rendered image in: " << output << std::endl; // main.cpp:688
        } else {
            std::cout << "Nothing saved: use -o to save rendered image" << std::endl;
        }
        tracer_scene_host_free(sorted_scene);
        tracer_scene_host_free(scene);
        tracer_cuda_shutdown();
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
}
