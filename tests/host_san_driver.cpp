// host_san_driver.cpp — test infrastructure: drives the product's HOST code (csrc/host_io.cpp, csrc/tracer_host.cpp:
// OBJ/MTL loader, flatten+sort, PPM writer, camera, mt19937 replay, band maths) under AddressSanitizer + UBSan.
// Built and run by tests/test_sanitizers_cpu.py; not part of the library.
//
//   host_san_driver <scratch dir> [model.obj ...]
//
// Every model is loaded (a refusal is fine: the reference refuses some of its own models, sceneloader.cpp:27-30),
// flattened + sorted, mapped back, and a small frame is written as P3 and P6.  Then the same entry points are fed
// hostile input: truncated statements, indices far outside the vertex list, empty files, missing materials.
// Exit code 0 = every call returned (with a status); any sanitizer report aborts the process.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/tracer_host.h"

static int n_loaded = 0, n_refused = 0;

static void exercise(const char *path, const std::string &scratch) {
    tracer_scene_host *h = nullptr;
    const int rc = tracer_scene_load_obj(path, &h);
    if (rc != TRACER_OK) {
        ++n_refused;
        if (h) std::abort(); // a refused scene must not hand out an object
        if (!tracer_host_last_error()[0]) std::abort(); // ... and must say why
        return;
    }
    ++n_loaded;
    const tracer_scene_flat *f = tracer_scene_host_flat(h);
    // read every array end to end (ASan checks the extents bind() promised)
    double sum = 0;
    const int N = f->n_geoms ? f->geom_tri_offset[f->n_geoms] : 0;
    for (int i = 0; i < 9 * N; ++i) sum += f->tri_verts[i];
    if (f->tri_normals)
        for (int i = 0; i < 9 * N; ++i) sum += f->tri_normals[i];
    for (int g = 0; g < f->n_geoms; ++g) {
        sum += f->geom_has_normals[g];
        for (int k = 0; k < 13; ++k) sum += f->geom_material[13 * g + k];
    }
    for (int l = 0; l < f->n_lights; ++l) sum += f->light_geom[l];
    tracer_scene_host *s = nullptr;
    if (tracer_scene_flatten_sorted(f, &s) == TRACER_OK) {
        const tracer_scene_flat *fs = tracer_scene_host_flat(s);
        const int NS = fs->n_geoms ? fs->geom_tri_offset[fs->n_geoms] : 0;
        for (int i = 0; i < 9 * NS; ++i) sum += fs->tri_verts[i];
        const int32_t *og = nullptr, *op = nullptr;
        if (NS > 0 && tracer_scene_host_origin(s, &og, &op) == TRACER_OK)
            for (int i = 0; i < NS; ++i) sum += og[i] + op[i];
        for (int i = 1; i < N; ++i) // the first N triangles are the sorted ones
            if (fs->tri_verts[9 * (size_t)i] < fs->tri_verts[9 * (size_t)(i - 1)]) std::abort();
        tracer_scene_host_free(s);
    }
    tracer_scene_host_free(h);
    if (sum != sum) std::printf("(nan in %s)\n", path);
}

static std::string put(const std::string &dir, const char *name, const std::string &text) {
    const std::string p = dir + "/" + name;
    FILE *f = std::fopen(p.c_str(), "wb");
    if (!f) std::abort();
    std::fwrite(text.data(), 1, text.size(), f);
    std::fclose(f);
    return p;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const std::string scratch = argv[1];
    for (int i = 2; i < argc; ++i) exercise(argv[i], scratch);

    // ---- hostile OBJ / MTL text --------------------------------------------------------------------------------
    put(scratch, "ok.mtl", "newmtl a\nKa 0.1 0.2 0.3\nKd 1 1 1\nNs 5\nnewmtl light\nKe 1 1 1\nKd\nKs 1\nNs\n");
    const char *hdr = "mtllib ok.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nusemtl a\n";
    const std::vector<std::string> bodies = {
        "",                                                   // no faces
        "f 1 2 3\n",                                          // plain
        "f 1 2\n",                                            // degenerate polygon: no triangle
        "f\n",                                                // empty face statement
        "f 1 2 3 4 5 6 7 8 9 10 11 12\n",                     // indices past the end
        "f -1 -2 -3\nf -4 -5 -6\n",                           // relative indices before the start
        "f 2147483647 1 2\nf -2147483648 1 2\n",              // extreme indices
        "f 99999999999999999999 1 2\n",                       // does not fit an int
        "f 1//1 2//7 3//-9\n",                                // normal indices out of range
        "f 1/1/1 2/2/2 3/3/3\n",                              // vt without any vt statement
        "f 1/ 2/ 3/\nf / / /\nf // // //\nf 1// 2// 3//\n",   // empty fields
        "f 0 0 0\n",                                          // index 0
        "v\nv 1\nv 1 2\nvn\nvn 1\nf 1 2 3\n",                 // short vertex statements
        "v 1e999 -1e999 nan\nv inf 0x10 1e\nv 1e+ .e5 +-1\nf 4 5 6\n",  // odd numbers
        "f 1 2 3\r\nf 1 2 3 \r\nf 1 2 3\t\r\n",               // CR / blank / tab line ends
        "g\no\ng a b c\nusemtl\nusemtl nothere\nf 1 2 3\n",   // unknown material -> refused, not crashed
        "usemtl light\nf 1 2 3\ng second\nusemtl a\nf 1 2 3\nusemtl light\nf 3 2 1\n",
        std::string("f 1 2 3") + std::string(100000, ' ') + "\n",       // a very long line
        std::string(50000, 'f') + "\n" + std::string("f ") + std::string(50000, '1') + " 2 3\n",
        std::string("f 1 2 3\0 4 5", 12) + "\n",              // an embedded NUL
        "mtllib\nmtllib  \nmtllib nothere.mtl\nf 1 2 3\n",    // bad mtllib statements: warnings, fatal
    };
    int k = 0;
    for (const std::string &b : bodies) {
        const std::string p = put(scratch, ("case" + std::to_string(k++) + ".obj").c_str(), hdr + b);
        exercise(p.c_str(), scratch);
    }
    exercise(put(scratch, "empty.obj", "").c_str(), scratch);
    exercise(put(scratch, "nolf.obj", "mtllib ok.mtl\nusemtl a\nv 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3").c_str(), scratch);
    exercise(put(scratch, "nomtl.obj", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n").c_str(), scratch);
    exercise((scratch + "/does_not_exist.obj").c_str(), scratch);
    put(scratch, "both.mtl", "newmtl a\nd 1\nTr 0\nnewmtl\nKd 1 1 1\n");
    exercise(put(scratch, "both.obj", "mtllib both.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl a\nf 1 2 3\n").c_str(), scratch);
    {
        tracer_scene_host *h = nullptr;
        if (tracer_scene_load_obj(nullptr, &h) == TRACER_OK) std::abort();
        if (tracer_scene_load_obj("x", nullptr) == TRACER_OK) std::abort();
        if (tracer_scene_flatten_sorted(nullptr, &h) == TRACER_OK) std::abort();
        if (tracer_scene_host_origin(nullptr, nullptr, nullptr) == TRACER_OK) std::abort();
        tracer_scene_host_free(nullptr);
        if (tracer_scene_host_flat(nullptr)) std::abort();
    }

    // ---- PPM writer --------------------------------------------------------------------------------------------------
    {
        const int W = 37, H = 23;
        std::vector<uint8_t> rgb((size_t)W * H * 3);
        for (size_t i = 0; i < rgb.size(); ++i) rgb[i] = (uint8_t)(i * 7);
        if (tracer_write_ppm((scratch + "/a.ppm").c_str(), rgb.data(), W, H, 0) != TRACER_OK) std::abort();
        if (tracer_write_ppm((scratch + "/b.ppm").c_str(), rgb.data(), W, H, 1) != TRACER_OK) std::abort();
        if (tracer_write_ppm((scratch + "/no/such/dir/c.ppm").c_str(), rgb.data(), W, H, 0) == TRACER_OK) std::abort();
        if (tracer_write_ppm(nullptr, rgb.data(), W, H, 0) == TRACER_OK) std::abort();
        if (tracer_write_ppm((scratch + "/d.ppm").c_str(), rgb.data(), 0, H, 0) == TRACER_OK) std::abort();
        // a frame larger than the writer's 1 MiB text buffer
        std::vector<uint8_t> big((size_t)640 * 360 * 3, 255);
        if (tracer_write_ppm((scratch + "/e.ppm").c_str(), big.data(), 640, 360, 0) != TRACER_OK) std::abort();
    }

    // ---- camera, RNG replay, band maths (csrc/tracer_host.cpp) ---------------------------------------------------
    {
        const float eye[3] = {0, 1, 2}, look[3] = {0, 1, 0}, up[3] = {0, 1, 0};
        tracer_camera cam;
        tracer_camera_lookat(eye, look, up, 60.f, 4.f / 3.f, &cam);
        tracer_camera_lookat(eye, eye, up, 60.f, 1.f, &cam); // degenerate: NaNs, not a crash
        const int W = 9, H = 7, L = 3;
        const int32_t off[5] = {0, 2, 3, 1000006, 1000007}, lights[L] = {0, 1, 2}; // lights of 2, 1 and 1000003 faces
        tracer_scene_flat sc{};
        sc.n_geoms = 4, sc.geom_tri_offset = off, sc.n_lights = L, sc.light_geom = lights;
        std::vector<uint8_t> hit((size_t)W * H);
        for (size_t i = 0; i < hit.size(); ++i) hit[i] = (i % 3) != 0;
        std::vector<int32_t> faceid((size_t)W * H * L, -7);
        if (tracer_mt19937_faceids(&sc, W, H, 1u, hit.data(), faceid.data()) != TRACER_OK) std::abort();
        for (size_t i = 0; i < faceid.size(); ++i) {
            const int F = off[i % L + 1] - off[i % L];
            if (hit[i / L] ? (faceid[i] < 0 || faceid[i] >= F) : faceid[i] != -1) std::abort();
        }
        if (tracer_mt19937_faceids(nullptr, W, H, 1u, hit.data(), faceid.data()) == TRACER_OK) std::abort();
        for (int rows = 1; rows <= 9; ++rows)
            for (int n = 1; n <= 8; ++n) {
                int total = 0;
                for (int b = 0; b < n; ++b) total += tracer_band_row_count(H, rows, b, n);
                if (total != H) std::abort();
            }
    }
    std::printf("host_san_driver: %d scenes loaded, %d refused, no sanitizer report\n", n_loaded, n_refused);
    return 0;
}
