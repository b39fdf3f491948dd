// host_io.cpp — OBJ/MTL loader and PPM writer (include/tracer_host.h).  Written from the
// behaviour of the reference's loader (src/scene/sceneloader.cpp over tinyobjloader v1.0.5),
// not from its code: a line tokenizer over std::string_view, explicit shape/material state.
#include <cctype>
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <string_view>
#include <vector>

#include "../../include/tracer_host.h"

struct tracer_scene_host {
    std::vector<int32_t> geom_tri_offset{0}, geom_has_normals, light_geom;
    std::vector<float> tri_verts, tri_normals, geom_material;
    std::vector<int32_t> origin_geom, origin_prim; // flatten+sort only: where each triangle came from
    tracer_scene_flat flat{};
    void bind() {
        flat.n_geoms = (int32_t)geom_has_normals.size();
        flat.geom_tri_offset = geom_tri_offset.data();
        flat.tri_verts = tri_verts.data();
        flat.tri_normals = tri_normals.empty() ? nullptr : tri_normals.data();
        flat.geom_has_normals = geom_has_normals.data();
        flat.geom_material = geom_material.data();
        flat.n_lights = (int32_t)light_geom.size();
        flat.light_geom = light_geom.data();
    }
};

namespace {

thread_local std::string g_host_err;

// ---- number parsing ---------------------------------------------------------------------
// Same conversion as tinyobj's tryParseDouble (tiny_obj_loader.h:465-580), which is NOT strtod:
// the mantissa is accumulated digit by digit in double (integer part m = m*10 + d; k-th fraction
// digit adds d * 10^-k with 10^-k taken from a literal table for k < 8 and pow(10,-k) beyond), and a
// decimal exponent e is applied as ldexp(m * pow(5,e), e).  Anything else would change the last bit
// of some vertex coordinates, and with it self-shadowing decisions.
bool parse_real(std::string_view tok, double &out) {
    size_t i = 0;
    const size_t n = tok.size();
    if (n == 0) return false;
    double sign = 1.0;
    if (tok[i] == '+' || tok[i] == '-') {
        sign = tok[i] == '-' ? -1.0 : 1.0;
        ++i;
    } else if (!std::isdigit((unsigned char)tok[i])) {
        return false;
    }
    double m = 0.0;
    size_t digits = 0;
    while (i < n && std::isdigit((unsigned char)tok[i])) {
        m *= 10;
        m += (int)(tok[i] - '0');
        ++i, ++digits;
    }
    if (digits == 0) return false;
    int e10 = 0;
    bool has_exp = false;
    if (i < n && tok[i] == '.') {
        static const double neg_pow10[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
        ++i;
        int k = 1;
        while (i < n && std::isdigit((unsigned char)tok[i])) {
            m += (int)(tok[i] - '0') * (k < 8 ? neg_pow10[k] : std::pow(10.0, -k));
            ++i, ++k;
        }
    } else if (i < n && (tok[i] == 'e' || tok[i] == 'E')) {
        // exponent directly after the integer part
    } else {
        out = sign * m; // trailing junk after the integer part is ignored
        return true;
    }
    if (i < n && (tok[i] == 'e' || tok[i] == 'E')) {
        ++i;
        int esign = 1;
        if (i < n && (tok[i] == '+' || tok[i] == '-')) {
            esign = tok[i] == '-' ? -1 : 1;
            ++i;
        } else if (!(i < n && std::isdigit((unsigned char)tok[i]))) {
            return false; // "1e" is a parse failure
        }
        size_t ed = 0;
        while (i < n && std::isdigit((unsigned char)tok[i])) {
            e10 = e10 * 10 + (int)(tok[i] - '0');
            ++i, ++ed;
        }
        if (ed == 0) return false;
        e10 *= esign;
        has_exp = true;
    }
    (void)has_exp;
    out = sign * (e10 ? std::ldexp(m * std::pow(5.0, e10), e10) : m);
    return true;
}

// A cursor over one line: tokens are separated by blanks/tabs; '\r' also ends a token.
struct Cursor {
    std::string_view s;
    size_t pos = 0;
    void skip_blank() {
        while (pos < s.size() && (s[pos] == ' ' || s[pos] == '\t')) ++pos;
    }
    std::string_view token() { // next blank-delimited token (may be empty at end of line)
        skip_blank();
        const size_t b = pos;
        while (pos < s.size() && s[pos] != ' ' && s[pos] != '\t' && s[pos] != '\r') ++pos;
        return s.substr(b, pos - b);
    }
    float real(double dflt = 0.0) {
        double v = dflt;
        const std::string_view t = token();
        double parsed;
        if (parse_real(t, parsed)) v = parsed;
        return (float)v;
    }
    bool at_end() {
        return pos >= s.size() || s[pos] == '\r' || s[pos] == '\n' || s[pos] == '\0';
    }
};

bool starts_with_key(std::string_view line, const char *key) { // key followed by a blank or tab
    const size_t k = std::strlen(key);
    return line.size() > k && line.compare(0, k, key) == 0 && (line[k] == ' ' || line[k] == '\t');
}

std::string first_word(std::string_view sv) { // what sscanf("%s") would read
    size_t b = 0;
    while (b < sv.size() && std::isspace((unsigned char)sv[b])) ++b;
    size_t e = b;
    while (e < sv.size() && !std::isspace((unsigned char)sv[e])) ++e;
    return std::string(sv.substr(b, e - b));
}

// reads a line ending in \n, \r\n or \r (or EOF); returns false when nothing is left
bool next_line(std::istream &is, std::string &line) {
    line.clear();
    if (is.peek() == EOF) return false;
    for (;;) {
        const int c = is.get();
        if (c == EOF || c == '\n') break;
        if (c == '\r') {
            if (is.peek() == '\n') is.get();
            break;
        }
        line.push_back((char)c);
    }
    return true;
}

struct Material {
    std::string name;
    float ka[3] = {0, 0, 0}, kd[3] = {0, 0, 0}, ks[3] = {0, 0, 0}, ke[3] = {0, 0, 0};
    float ns = 1.f; // tinyobj's default shininess (tiny_obj_loader.h:851)
};

// MTL: the statements the render path consumes (Ka Kd Ks Ke Ns) plus the one that can raise a
// warning (d together with Tr).  Returns warnings (non-empty => the reference would refuse the scene).
std::string load_mtl(std::istream &is, std::vector<Material> &mats, std::map<std::string, int> &by_name) {
    std::string warnings, line;
    Material cur;
    bool has_d = false, has_tr = false;
    auto flush = [&]() {
        by_name.insert({cur.name, (int)mats.size()}); // first definition of a name wins (std::map::insert)
        mats.push_back(cur);
    };
    while (next_line(is, line)) {
        const size_t last = line.find_last_not_of(" \t");
        line = last == std::string::npos ? std::string() : line.substr(0, last + 1);
        Cursor c{line};
        c.skip_blank();
        const std::string_view rest = std::string_view(line).substr(c.pos);
        if (rest.empty() || rest[0] == '#') continue;
        if (starts_with_key(rest, "newmtl")) {
            if (!cur.name.empty()) flush();
            cur = Material();
            has_d = has_tr = false;
            cur.name = first_word(rest.substr(7));
            continue;
        }
        auto rgb = [&](float *dst) {
            Cursor v{rest, 2};
            dst[0] = v.real(), dst[1] = v.real(), dst[2] = v.real();
        };
        if (starts_with_key(rest, "Ka")) rgb(cur.ka);
        else if (starts_with_key(rest, "Kd")) rgb(cur.kd);
        else if (starts_with_key(rest, "Ks")) rgb(cur.ks);
        else if (starts_with_key(rest, "Ke")) rgb(cur.ke);
        else if (starts_with_key(rest, "Ns")) {
            Cursor v{rest, 2};
            cur.ns = v.real();
        } else if (starts_with_key(rest, "d")) {
            if (has_tr) warnings += "WARN: Both `d` and `Tr` parameters defined for \"" + cur.name + "\".\n";
            has_d = true;
        } else if (starts_with_key(rest, "Tr")) {
            if (has_d) warnings += "WARN: Both `d` and `Tr` parameters defined for \"" + cur.name + "\".\n";
            has_tr = true;
        }
    }
    flush(); // the last material is stored even when it has no name
    return warnings;
}

struct Corner {
    int v = -1, vt = -1, vn = -1;
};

int resolve_index(int idx, int count) { // OBJ indices: 1-based, negative = relative to the current count
    if (idx > 0) return idx - 1;
    if (idx == 0) return 0;
    return count + idx;
}

Corner parse_corner(std::string_view tok, int nv, int nvn, int nvt) {
    Corner c;
    const std::string s(tok);
    const char *p = s.c_str();
    c.v = resolve_index(std::atoi(p), nv);
    p += std::strcspn(p, "/");
    if (*p != '/') return c;
    ++p;
    if (*p == '/') { // i//k
        ++p;
        c.vn = resolve_index(std::atoi(p), nvn);
        return c;
    }
    c.vt = resolve_index(std::atoi(p), nvt);
    p += std::strcspn(p, "/");
    if (*p != '/') return c;
    ++p;
    c.vn = resolve_index(std::atoi(p), nvn);
    return c;
}

struct Shape { // triangulated faces of one OBJ shape; material id per triangle
    std::vector<Corner> corners;
    std::vector<int> material_ids;
};

// move the pending polygons into the shape as a triangle fan; false when there was nothing pending
bool flush_faces(Shape &shape, std::vector<std::vector<Corner>> &pending, int material) {
    if (pending.empty()) return false;
    for (const auto &poly : pending) {
        for (size_t k = 2; k < poly.size(); ++k) {
            shape.corners.push_back(poly[0]);
            shape.corners.push_back(poly[k - 1]);
            shape.corners.push_back(poly[k]);
            shape.material_ids.push_back(material);
        }
    }
    return true;
}

}  // namespace

extern "C" {

const char *tracer_host_last_error(void) { return g_host_err.c_str(); }

int tracer_scene_load_obj(const char *obj_path, tracer_scene_host **out) {
    if (!out) return TRACER_ERR_INVALID;
    *out = nullptr;
    if (!obj_path) {
        g_host_err = "null path";
        return TRACER_ERR_INVALID;
    }
    const std::string path(obj_path);
    std::ifstream in(path);
    if (!in) {
        g_host_err = "Cannot open file [" + path + "]";
        return TRACER_ERR_INVALID;
    }
    const std::string base_dir = path.substr(0, path.rfind('/') + 1); // sceneloader.cpp:20

    std::vector<float> v, vn;
    int n_vt = 0;
    std::vector<Material> mats;
    std::map<std::string, int> mat_by_name;
    std::vector<Shape> shapes;
    Shape shape;
    std::vector<std::vector<Corner>> pending;
    int material = -1;
    std::string warnings, line;

    while (next_line(in, line)) {
        Cursor c{line};
        c.skip_blank();
        const std::string_view rest = std::string_view(line).substr(c.pos);
        if (rest.empty() || rest[0] == '#') continue;
        if (starts_with_key(rest, "v")) {
            Cursor p{rest, 2};
            for (int k = 0; k < 3; ++k) v.push_back(p.real());
        } else if (starts_with_key(rest, "vn")) {
            Cursor p{rest, 3};
            for (int k = 0; k < 3; ++k) vn.push_back(p.real());
        } else if (starts_with_key(rest, "vt")) {
            ++n_vt;
        } else if (starts_with_key(rest, "f")) {
            Cursor p{rest, 2};
            std::vector<Corner> poly;
            for (;;) {
                p.skip_blank();
                if (p.at_end()) break;
                poly.push_back(parse_corner(p.token(), (int)v.size() / 3, (int)vn.size() / 3, n_vt));
                while (p.pos < rest.size() && rest[p.pos] == '\r') ++p.pos;
            }
            pending.push_back(std::move(poly));
        } else if (starts_with_key(rest, "usemtl")) {
            const auto it = mat_by_name.find(first_word(rest.substr(7)));
            const int id = it == mat_by_name.end() ? -1 : it->second;
            if (id != material) { // the shape keeps growing; only the pending faces take the old material
                flush_faces(shape, pending, material);
                pending.clear();
                material = id;
            }
        } else if (starts_with_key(rest, "mtllib")) {
            std::vector<std::string> files;
            {
                std::stringstream ss{std::string(rest.substr(7))};
                std::string item;
                while (std::getline(ss, item, ' ')) files.push_back(item);
            }
            if (files.empty()) {
                warnings += "WARN: Looks like empty filename for mtllib.\n";
            } else {
                bool found = false;
                for (const auto &f : files) {
                    std::ifstream mtl(base_dir + f);
                    if (!mtl) {
                        warnings += "WARN: Material file [ " + base_dir + f + " ] not found.\n";
                        continue;
                    }
                    warnings += load_mtl(mtl, mats, mat_by_name);
                    found = true;
                    break;
                }
                if (!found) warnings += "WARN: Failed to load material file(s).\n";
            }
        } else if (starts_with_key(rest, "g") || starts_with_key(rest, "o")) {
            // a new group/object closes the current shape — but only if faces are pending: faces already
            // moved into the shape by a usemtl are dropped with it otherwise (tiny_obj_loader.h:1591-1648)
            if (flush_faces(shape, pending, material)) shapes.push_back(shape);
            shape = Shape();
            pending.clear();
        }
    }
    if (flush_faces(shape, pending, material) || !shape.corners.empty()) shapes.push_back(shape);

    if (!warnings.empty()) { // sceneloader.cpp:27-30: warnings are fatal
        g_host_err = "TinyOBJ Error loading " + path + " error: " + warnings;
        return TRACER_ERR_INVALID;
    }

    auto *h = new tracer_scene_host();
    for (const Shape &sh : shapes) {
        if (sh.material_ids.empty() || sh.material_ids[0] < 0 || sh.material_ids[0] >= (int)mats.size()) {
            g_host_err = "shape without a valid material (the reference indexes obj_materials[material_ids[0]], sceneloader.cpp:52)";
            delete h;
            return TRACER_ERR_INVALID;
        }
        const Material &m = mats[sh.material_ids[0]];
        const float *src[4] = {m.ka, m.kd, m.ks, m.ke};
        for (const float *q : src) h->geom_material.insert(h->geom_material.end(), q, q + 3);
        h->geom_material.push_back(m.ns);
        float ke2 = 0; // dot(ke, ke) > 0, float accumulation order of vec.h:95-101
        ke2 += m.ke[0] * m.ke[0];
        ke2 += m.ke[1] * m.ke[1];
        ke2 += m.ke[2] * m.ke[2];
        bool any_normal = false;
        const size_t first_normal = h->tri_normals.size();
        for (const Corner &c : sh.corners) {
            if (c.v < 0 || (size_t)c.v * 3 + 2 >= v.size()) {
                g_host_err = "face references a vertex that does not exist";
                delete h;
                return TRACER_ERR_INVALID;
            }
            for (int k = 0; k < 3; ++k) h->tri_verts.push_back(v[(size_t)c.v * 3 + k]);
            if (c.vn != -1 && c.vn >= 0 && (size_t)c.vn * 3 + 2 < vn.size()) {
                const float nx = vn[(size_t)c.vn * 3], ny = vn[(size_t)c.vn * 3 + 1], nz = vn[(size_t)c.vn * 3 + 2];
                float d = 0; // normalize(n) = n / sqrt(dot(n,n)) (vec.h:135-137)
                d += nx * nx;
                d += ny * ny;
                d += nz * nz;
                const float l = std::sqrt(d);
                h->tri_normals.insert(h->tri_normals.end(), {nx / l, ny / l, nz / l});
                any_normal = true;
            } else {
                h->tri_normals.insert(h->tri_normals.end(), {0.f, 0.f, 0.f});
            }
        }
        (void)first_normal;
        h->geom_has_normals.push_back(any_normal ? 1 : 0);
        h->geom_tri_offset.push_back(h->geom_tri_offset.back() + (int32_t)(sh.corners.size() / 3));
        if (ke2 > 0) h->light_geom.push_back((int32_t)h->geom_has_normals.size() - 1);
    }
    h->bind();
    *out = h;
    return TRACER_OK;
}

const tracer_scene_flat *tracer_scene_host_flat(const tracer_scene_host *scene) { return scene ? &scene->flat : nullptr; }

// flatten + sort, src/simplify/flatten.cpp:50-82 with the comparator of flatten.cpp:20-27 (see tracer_host.h)
int tracer_scene_flatten_sorted(const tracer_scene_flat *in, tracer_scene_host **out) {
    if (!out) return TRACER_ERR_INVALID;
    *out = nullptr;
    if (!in || in->n_geoms < 0 || (in->n_geoms > 0 && (!in->geom_tri_offset || !in->geom_material))) {
        g_host_err = "flatten_sorted: bad scene";
        return TRACER_ERR_INVALID;
    }
    if (in->n_spheres > 0) {
        g_host_err = "flatten_sorted: the reference's flatten pass knows triangles only";
        return TRACER_ERR_INVALID;
    }
    const int G = in->n_geoms;
    const int N = G > 0 ? in->geom_tri_offset[G] : 0;
    std::vector<int32_t> tri_geom((size_t)N), order((size_t)N);
    for (int g = 0; g < G; ++g)
        for (int t = in->geom_tri_offset[g]; t < in->geom_tri_offset[g + 1]; ++t) tri_geom[t] = g;
    for (int t = 0; t < N; ++t) order[t] = t; // flatten.cpp:53-75: geometry major, face minor
    // flatten.cpp:78 sorts with leftMostTriangle (vertices[0].x ascending, flatten.cpp:20-27).  std::sort leaves the
    // order of equal keys unspecified; a stable sort is one of its valid outcomes and makes the order reproducible.
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return in->tri_verts[9 * (size_t)a] < in->tri_verts[9 * (size_t)b]; });
    auto *h = new tracer_scene_host();
    bool any_normals = false;
    for (int g = 0; g < G && in->geom_has_normals; ++g) any_normals |= in->geom_has_normals[g] != 0;
    any_normals = any_normals && in->tri_normals;
    auto open_geom = [&](int g) {
        h->geom_has_normals.push_back(in->geom_has_normals ? in->geom_has_normals[g] : 0);
        h->geom_material.insert(h->geom_material.end(), in->geom_material + 13 * (size_t)g, in->geom_material + 13 * (size_t)g + 13);
        h->geom_tri_offset.push_back(h->geom_tri_offset.back());
    };
    auto push_tri = [&](int t) {
        h->tri_verts.insert(h->tri_verts.end(), in->tri_verts + 9 * (size_t)t, in->tri_verts + 9 * (size_t)t + 9);
        if (any_normals) h->tri_normals.insert(h->tri_normals.end(), in->tri_normals + 9 * (size_t)t, in->tri_normals + 9 * (size_t)t + 9);
        h->origin_geom.push_back(tri_geom[t]);
        h->origin_prim.push_back(t - in->geom_tri_offset[tri_geom[t]]); // c_triangle.prim_id, flatten.cpp:63
        ++h->geom_tri_offset.back();
    };
    // maximal runs of one original geometry become geometries of the output: iteration order == sorted order,
    // material / has-normals of every triangle unchanged
    for (int i = 0; i < N; ++i) {
        const int t = order[i];
        if (i == 0 || tri_geom[t] != tri_geom[order[i - 1]]) open_geom(tri_geom[t]);
        push_tri(t);
    }
    // light geometries once more, at the end, in their ORIGINAL face order: main.cpp:749 samples light.vertex[faceID]
    // from the light's own vertex list, which sorting must not permute.  The copies cannot change any result: a copy
    // lies behind its original in the iteration order and computes the same t2, which `t2 >= t` rejects in the
    // closest-hit loop (ray_triangle.h:49), and an occlusion ray that would stop at a copy has already stopped at
    // the original (main.cpp:317-325).
    for (int l = 0; l < in->n_lights; ++l) {
        const int g = in->light_geom[l];
        if (g < 0 || g >= G) {
            delete h;
            g_host_err = "flatten_sorted: light_geom index out of range";
            return TRACER_ERR_INVALID;
        }
        open_geom(g);
        for (int t = in->geom_tri_offset[g]; t < in->geom_tri_offset[g + 1]; ++t) push_tri(t);
        h->light_geom.push_back((int32_t)h->geom_has_normals.size() - 1);
    }
    h->bind();
    *out = h;
    return TRACER_OK;
}

int tracer_scene_host_origin(const tracer_scene_host *scene, const int32_t **geom, const int32_t **prim) {
    if (!scene || scene->origin_geom.empty()) return TRACER_ERR_INVALID;
    if (geom) *geom = scene->origin_geom.data();
    if (prim) *prim = scene->origin_prim.data();
    return TRACER_OK;
}

void tracer_scene_host_free(tracer_scene_host *scene) { delete scene; }

int tracer_write_ppm(const char *path, const uint8_t *rgb, int32_t width, int32_t height, int32_t binary) {
    if (!path || !rgb || width <= 0 || height <= 0) return TRACER_ERR_INVALID;
    std::FILE *f = std::fopen(path, "wb");
    if (!f) {
        g_host_err = std::string("cannot open ") + path;
        return TRACER_ERR_INVALID;
    }
    const size_t n = (size_t)width * height;
    if (binary) {
        std::fprintf(f, "P6\n%d %d\n255\n", width, height);
        std::fwrite(rgb, 1, n * 3, f);
    } else {
        // P3 through a small-integer text table: 25-100 M integers at 4K/8K would otherwise dominate
        char lut[256][4];
        int len[256];
        for (int i = 0; i < 256; ++i) len[i] = std::snprintf(lut[i], 4, "%d", i);
        std::fprintf(f, "P3\n%d %d\n255\n", width, height);
        std::vector<char> buf;
        buf.reserve(1 << 20);
        for (size_t i = 0; i < n; ++i) {
            for (int c = 0; c < 3; ++c) {
                const uint8_t x = rgb[i * 3 + c];
                buf.insert(buf.end(), lut[x], lut[x] + len[x]);
                buf.push_back(c == 2 ? '\n' : ' ');
            }
            if (buf.size() > (1 << 20) - 16) {
                std::fwrite(buf.data(), 1, buf.size(), f);
                buf.clear();
            }
        }
        std::fwrite(buf.data(), 1, buf.size(), f);
    }
    std::fclose(f);
    return TRACER_OK;
}

}  // extern "C"
