#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Compile-proof of the drop-in.

Reads the reference's src/main.cpp WHERE IT LIES, applies the binding a maintainer would add — a `--cuda` flag beside
`--ispc`, the flat-scene marshalling, one call to tracer_cuda_render where ispc::trace is called (src/main.cpp:591-625)
and the packed bytes handed to the reference's own PPM writer — and writes the patched translation unit to the path
given (under oracle/_ref/, git-ignored: no reference source enters the repository).  oracle/Makefile compiles it with
-DUSE_CUDA against include/tracer_cuda.h and links it with libtracer_cuda.so into oracle/_ref/ESCViewer2021_cuda.
`--seed N` (also added) makes the run reproducible: the reference seeds from std::random_device (main.cpp:587-588).

    python oracle/apply_cuda_patch.py /root/reference/src/main.cpp oracle/_ref/main_cuda.cpp
"""
import sys

HELPER = r'''
#ifdef USE_CUDA
// ---- INTEGRATION.md section 3: the binding -------------------------------------------------------------------
#include "tracer_cuda.h"
static void flatten_scene_cuda(const tracer::scene &S, std::vector<int32_t> &off, std::vector<float> &verts,
                               std::vector<float> &normals, std::vector<int32_t> &has_n, std::vector<float> &mats,
                               std::vector<int32_t> &lights) {
    off.push_back(0);
    for (const auto &g : S.geometry) {                       // geometry major ...
        for (const auto &face : g.face_index)                // ... face minor (main.cpp:179-180)
            for (int c = 0; c < 3; ++c) {
                const auto &v = g.vertex[face[c]];
                verts.insert(verts.end(), {v.x, v.y, v.z});
                if (!g.normals.empty()) { const auto &n = g.normals[face[c]]; normals.insert(normals.end(), {n.x, n.y, n.z}); }
                else normals.insert(normals.end(), {0.f, 0.f, 0.f});
            }
        off.push_back(off.back() + (int32_t)g.face_index.size());
        has_n.push_back(!g.normals.empty());                 // main.cpp:733
        const auto &m = g.object_material;
        mats.insert(mats.end(), {m.ka.x, m.ka.y, m.ka.z, m.kd.x, m.kd.y, m.kd.z, m.ks.x, m.ks.y, m.ks.z,
                                 m.ke.x, m.ke.y, m.ke.z, m.Ns});
    }
    for (size_t l : S.light_sources) lights.push_back((int32_t)l);
}
#endif

'''

FLAGS = r'''
#ifdef USE_CUDA
        // --cuda renders on the GPU through libtracer_cuda (beside --ispc)
        if (std::string(argv[arg]) == "--cuda") {
            cuda = true;
            continue;
        }
        // --seed N : reproducible light sampling (the stock binary seeds from std::random_device)
        if (std::string(argv[arg]) == "--seed") {
            cuda_seed = (unsigned)std::strtoul(argv[arg + 1], nullptr, 10);
            have_seed = true;
            arg++;
            continue;
        }
#endif
'''

RENDER = r'''
#ifdef USE_CUDA
    if (cuda) {                                                   // where `if (ispc) {` is (main.cpp:591)
        std::vector<int32_t> off, has_n, lights; std::vector<float> verts, normals, mats;
        flatten_scene_cuda(SceneMesh, off, verts, normals, has_n, mats, lights);
        tracer_scene_flat fs{};
        fs.n_geoms = (int32_t)SceneMesh.geometry.size();  fs.geom_tri_offset = off.data();
        fs.tri_verts = verts.data();  fs.tri_normals = normals.data();  fs.geom_has_normals = has_n.data();
        fs.geom_material = mats.data();  fs.n_lights = (int32_t)lights.size();  fs.light_geom = lights.data();
        tracer_camera tc;  const float e[3] = {eye.x, eye.y, eye.z}, l[3] = {look.x, look.y, look.z}, up[3] = {0, 1, 0};
        tracer_camera_lookat(e, l, up, vfov, aspect, &tc);        // same arithmetic as main.cpp:551
        tracer_render_opts o{};  o.struct_size = sizeof o;  o.rng_mode = TRACER_RNG_MT19937;  o.seed = have_seed ? cuda_seed : rd();
        cuda_rgb.resize((size_t)image_width * image_height * 3);
        if (tracer_cuda_init(0) || tracer_cuda_render(&fs, &tc, image_width, image_height, &o, cuda_rgb.data()))
            throw std::runtime_error(tracer_cuda_last_error());
    }
    else
#endif
'''

WRITER = r'''
#ifdef USE_CUDA
                if (cuda) { // rows are already top-to-bottom and quantised (main.cpp:679-684 done on the GPU)
                    const size_t o3 = ((size_t)(image_height - 1 - h) * image_width + w) * 3;
                    file << int(cuda_rgb[o3]) << " " << int(cuda_rgb[o3 + 1]) << " " << int(cuda_rgb[o3 + 2]) << "\n";
                    continue;
                }
#endif
'''


def once(text, anchor, new, before=True):
    n = text.count(anchor)
    if n < 1:
        raise SystemExit(f"anchor not found in the reference's main.cpp: {anchor!r}")
    i = text.index(anchor)
    return text[:i] + (new + anchor if before else anchor + new) + text[i + len(anchor):]


def main(src, dst):
    t = open(src).read()
    t = once(t, "int main(int argc, char *argv[]) {", HELPER)
    t = once(t, "    bool ispc{false};\n", "    bool cuda{false}, have_seed{false};\n    unsigned cuda_seed{0};\n    std::vector<uint8_t> cuda_rgb;\n", before=False)
    t = once(t, "        // --test - only test ispc functions", FLAGS.lstrip("\n"))
    t = once(t, "    if (ispc) {\n        // prepare a flat array of floats for image pixels in ispc", RENDER.lstrip("\n"))
    t = once(t, "                tracer::vec3<float> img;\n", WRITER.lstrip("\n"))
    open(dst, "w").write(t)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
