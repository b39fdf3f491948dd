"""Fuzz of the OBJ/MTL loader (csrc/host_io.cpp) against the reference's own loader (model::loadobj over tinyobjloader
v1.0.5, through oracle/_ref): random OBJ text — triangles / quads / polygons, i, i/j, i//k, i/j/k and negative indices, g / o / usemtl
/ s statements in random places, comments, blank lines, CRLF, tabs, trailing blanks, numbers in many spellings — must load to the
same flat scene bit for bit, or be refused by both.  Needs /root/reference (test infrastructure).
    python tests/fuzz_loader.py [n_files] [seed]
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from esctp1raytracer_b200 import Scene, TracerError  # noqa: E402

MTL = """# materials
newmtl white
Ka 0 0 0
Kd 0.725 0.71 0.68
Ks 0 0 0
Ns 10

newmtl red
Kd 0.63 0.065 0.05
Ks 0.5 0.5 0.5
Ns 32.5
newmtl light
Ka 0 0 0
Kd 0.78 0.78 0.78
Ke 17 12 4
newmtl dim
\tKd\t0.1  0.2   0.3
Ke 0 0 0.001
illum 2
"""
MATS = ["white", "red", "light", "dim"]


ODD = ["0.1234567890123456789", "123456789012345678901234567890", "1e40", "-1e40", "1e-45", "4.9e-324", "1e-400", "1e400", "-0", "-0.0", ".5e1",
       "5e", "e5", "1.e2", "+.5", "--1", "1..2", "0x10", "nan", "inf", "-inf", "1e+", "1e-", ".", "-", "+", "1.5.5", "2,5", "1_000",
       "00012.5000", "1E5", "1e05", "1e+05", "3.4028235e38", "3.4028236e38", "1.17549435e-38", "16777217", "0.30000000000000004",
       "9007199254740993", "1e22", "1e23", "8.5e-1abc", "7junk", "1d5", "1f"]


def num(rng, x):
    if rng.integers(0, 12) == 0:
        return ODD[rng.integers(0, len(ODD))]
    k = rng.integers(0, 9)
    if k == 0:
        return "%g" % x
    if k == 1:
        return "%.6f" % x
    if k == 2:
        return "%e" % x
    if k == 3:
        return "%+.4f" % x
    if k == 4:
        return ("%.3f" % x).replace("0.", ".", 1) if abs(x) < 1 else "%.3f" % x
    if k == 5:
        return "%d" % round(x)
    if k == 6:
        return "%d." % round(x)
    if k == 7:
        return "%.9g" % x
    return "%.2E" % x


def make_mtl(rng):
    """a material library that defines MATS (possibly more than once: the first definition wins) in random spellings"""
    out = ["# fuzz mtl"]
    names = list(MATS) + [MATS[i] for i in rng.integers(0, len(MATS), int(rng.integers(0, 3)))]
    rng.shuffle(names)
    both = False
    for nm in names:
        out.append(rng.choice(["newmtl ", "newmtl\t", "  newmtl "]) + nm + rng.choice(["", " ", "  extra words"]))
        stmts = []
        for key in ("Ka", "Kd", "Ks", "Ke"):
            if rng.integers(0, 4) == 0:
                continue
            vals = rng.uniform(0, 1, int(rng.choice([1, 2, 3, 3, 3, 4]))) * (rng.choice([0, 1, 20]) if key == "Ke" else 1)
            stmts.append(rng.choice(["", " ", "\t"]) + key + rng.choice([" ", "\t", "  "]) + " ".join(num(rng, v) for v in vals))
        if rng.integers(0, 2):
            stmts.append("Ns " + num(rng, rng.uniform(0, 200)))
        stmts += [rng.choice(["illum 2", "Ni 1.5", "map_Kd tex.png", "# c", "", "Tf 1 1 1", "bump b.png", "Kx 1 2 3"]) for _ in range(int(rng.integers(0, 3)))]
        r = rng.integers(0, 12)
        if r == 0:
            stmts.append("d " + num(rng, 0.5))
        elif r == 1:
            stmts.append("Tr " + num(rng, 0.5))
        elif r == 2:  # both: a loader warning, fatal in the reference (sceneloader.cpp:27-30)
            stmts += ["d 0.5", "Tr 0.5"]
            both = True
        rng.shuffle(stmts)
        out += stmts
    return "\n".join(out) + rng.choice(["", "\n"]), both


def make_obj(rng):
    eol = "\r\n" if rng.integers(0, 4) == 0 else "\n"
    lines = ["# fuzz", rng.choice(["mtllib m.mtl", "mtllib m.mtl other.mtl", "mtllib missing.mtl m.mtl", "mtllib  m.mtl", "mtllib m.mtl "])]
    nv = nn = nt = 0
    used_mtl = False
    n_stmt = int(rng.integers(5, 120))
    # One index form per file.  A geometry that mixes faces with and without normals is outside the reference's own contract:
    # its loader packs only the normals that exist (sceneloader.cpp:84-89) and its renderer then indexes that shorter array
    # with vertex indices (main.cpp:733-737, out of bounds); ours keeps the arrays aligned (zeros where a corner has no
    # normal).  Likewise a face before any usemtl makes the reference index obj_materials[-1] (sceneloader.cpp:52): ours
    # refuses the file, the reference crashes or throws at random.
    file_form = int(rng.integers(0, 4))
    if file_form in (2, 3):
        for _ in range(int(rng.integers(1, 4))):
            lines.append("vn " + " ".join(num(rng, v) for v in rng.normal(size=3)))
            nn += 1
    if file_form in (1, 3):
        for _ in range(int(rng.integers(1, 4))):
            lines.append("vt " + " ".join(num(rng, v) for v in rng.uniform(size=2)))
            nt += 1
    for _ in range(n_stmt):
        k = rng.integers(0, 100)
        if k < 35 or nv < 3:
            xyz = rng.normal(size=3) * 10 ** rng.uniform(-2, 2)
            lines.append("v " + " ".join(num(rng, v) for v in xyz) + (" 1.0" if rng.integers(0, 10) == 0 else ""))
            nv += 1
        elif k < 42:
            lines.append("vn " + " ".join(num(rng, v) for v in rng.normal(size=3)))
            nn += 1
        elif k < 46:
            lines.append("vt " + " ".join(num(rng, v) for v in rng.uniform(size=2)))
            nt += 1
        elif k < 78:
            if not used_mtl:
                lines.append("usemtl " + MATS[rng.integers(0, len(MATS))])
                used_mtl = True
            cnt = int(rng.choice([3, 3, 3, 4, 4, 5, 6]))
            form = file_form
            toks = []
            for _c in range(cnt):
                neg = rng.integers(0, 4) == 0
                vi = -int(rng.integers(1, nv + 1)) if neg else int(rng.integers(1, nv + 1))
                ti = (-int(rng.integers(1, nt + 1)) if neg else int(rng.integers(1, nt + 1))) if nt else 1
                ni = (-int(rng.integers(1, nn + 1)) if neg else int(rng.integers(1, nn + 1))) if nn else 1
                toks.append([f"{vi}", f"{vi}/{ti}", f"{vi}//{ni}", f"{vi}/{ti}/{ni}"][form])
            sep = "\t" if rng.integers(0, 8) == 0 else " "
            lines.append("f" + sep + sep.join(toks) + (" " if rng.integers(0, 5) == 0 else ""))
        elif k < 84:
            lines.append(("g " if rng.integers(0, 2) else "o ") + "part%d" % rng.integers(0, 50) + (" extra" if rng.integers(0, 5) == 0 else ""))
        elif k < 92:
            lines.append("usemtl " + MATS[rng.integers(0, len(MATS))])
            used_mtl = True
        elif k < 95:
            lines.append("s " + rng.choice(["off", "1", "2"]))
        elif k < 98:
            lines.append(rng.choice(["", "   ", "# a comment", "#", "\t# indented comment"]))
        else:
            lines.append("  " + "v " + " ".join(num(rng, v) for v in rng.normal(size=3)))  # indented statement
            nv += 1
    text = eol.join(lines)
    if rng.integers(0, 3):
        text += eol
    return text


def same(a, b):
    bits = lambda x: np.ascontiguousarray(x, np.float32).view(np.uint32)
    if not (np.array_equal(a.geom_tri_offset, b.geom_tri_offset) and np.array_equal(bits(a.tri_verts), bits(b.tri_verts))):
        return "geometry"
    if not np.array_equal(a.geom_has_normals, b.geom_has_normals):
        return "has_normals"
    if not np.array_equal(bits(a.geom_material), bits(b.geom_material)):
        return "materials"
    if not np.array_equal(a.light_geom, b.light_geom):
        return "lights"
    if b.tri_normals is not None and not np.array_equal(bits(a.tri_normals), bits(b.tri_normals)):
        return "normals"
    return None


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    ref = oracle.RefOracle()
    n_ok = n_both_refuse = n_bad = n_ref_crash = n_ref_ub = n_crash_ours_loaded = 0
    with tempfile.TemporaryDirectory() as d:
        for i in range(n):
            path = os.path.join(d, "f.obj")
            mtl = MTL if i % 3 == 0 else make_mtl(rng)[0]
            open(os.path.join(d, "m.mtl"), "w", newline="").write(mtl)
            text = make_obj(rng)
            open(path, "w", newline="").write(text)
            want = got = None
            # the reference loader runs in a forked child: on input outside its contract it crashes instead of refusing
            dump = os.path.join(d, "want.npz")
            pid = os.fork()
            if pid == 0:
                code = 1
                try:
                    h = ref.load_obj(path)
                    w = ref.dump(h)
                    np.savez(dump, off=w.geom_tri_offset, tv=w.tri_verts, hn=w.geom_has_normals, mat=w.geom_material, lg=w.light_geom,
                             tn=w.tri_normals if w.tri_normals is not None else np.zeros(0, np.float32))
                    code = 0
                except RuntimeError as e:
                    open(dump + ".err", "w").write(str(e))
                    code = 3
                finally:
                    os._exit(code)
            _, status = os.waitpid(pid, 0)
            if os.WIFSIGNALED(status):
                n_ref_crash += 1
                try:
                    Scene.load_obj(path)  # ours must still return on the same input ...
                    n_crash_ours_loaded += 1
                    keep = os.path.join(tempfile.gettempdir(), f"fuzz_loader_refcrash_{i}.obj")
                    open(keep, "w", newline="").write(text.replace("m.mtl", f"fuzz_loader_refcrash_{i}.mtl"))
                    open(keep[:-4] + ".mtl", "w", newline="").write(mtl)
                except TracerError:
                    pass  # ... normally by refusing it
                continue
            if os.WEXITSTATUS(status) == 0:
                z = np.load(dump)
                want = oracle.FlatScene(z["off"], z["tv"], z["tn"] if len(z["tn"]) else None, z["hn"], z["mat"], z["lg"])
            else:
                werr = open(dump + ".err").read() if os.path.exists(dump + ".err") else "?"
            try:
                got = Scene.load_obj(path)
            except TracerError as e:
                gerr = str(e)
            if want is None and got is None:
                n_both_refuse += 1
                continue
            if got is None and "shape without a valid material" in gerr:
                # the reference indexes obj_materials[material_ids[0]] with id -1 here (sceneloader.cpp:52): undefined behaviour
                # that happens not to crash — e.g. `mtllib  m.mtl` with two blanks makes tinyobj "load" the directory first
                # (tiny_obj_loader.h:1554-1575) and never reach the real file.  Ours refuses the file.
                n_ref_ub += 1
                continue
            why = "one side refused" if (want is None) != (got is None) else same(got, want)
            if why:
                n_bad += 1
                keep = os.path.join(tempfile.gettempdir(), f"fuzz_loader_bad_{i}.obj")
                open(keep, "w", newline="").write(text.replace("m.mtl", f"fuzz_loader_bad_{i}.mtl"))
                open(keep[:-4] + ".mtl", "w", newline="").write(mtl)
                print(f"MISMATCH file {i}: {why}  (kept {keep}; ref {'refused: ' + werr[:80] if want is None else 'ok'}, ours {'refused: ' + gerr[:80] if got is None else 'ok'})")
            else:
                n_ok += 1
    print(f"fuzz_loader: {n} files: {n_ok} equal bit for bit, {n_both_refuse} refused by both, {n_ref_crash} crashed the reference's loader "
          f"and {n_ref_ub} made it index obj_materials[-1] without crashing (ours refused these; "
          f"{n_crash_ours_loaded} of the crashing files loaded here), {n_bad} mismatches")
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
