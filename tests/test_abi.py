"""CPU tests of the drop-in boundary: libflgpu.so loads without a GPU, exports every symbol that
include/flgpu.h declares, its structs have the layout the ctypes mirror assumes, C++ programs written
against the reference-style wrappers link against it, and it refuses to compute without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "flgpu.h")
LIB = os.path.join(ROOT, "fortran_library_b200", "libflgpu.so")


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "fortran_library_b200", "csrc"), "../libflgpu.so"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b((?:flgpu_|__nonlinearoptimization_MOD_|nonlinearoptimization_mp_)\w+)\s*\(", src))
    # function-pointer typedefs look like (*flgpu_xxx_fn)( -- not functions
    typedefs = set(re.findall(r"\(\*\s*(\w+)\s*\)", src))
    return sorted(names - typedefs)


def test_library_exports_every_declared_symbol():
    _build()
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    declared = _declared_functions()
    assert len(declared) >= 30
    missing = [d for d in declared if d not in exported]
    assert not missing, f"declared in flgpu.h but not exported: {missing}"
    for sym in ("__nonlinearoptimization_MOD_lbfgs", "__nonlinearoptimization_MOD_conjugategradient",
                "__nonlinearoptimization_MOD_conjugategradient_basic", "nonlinearoptimization_mp_lbfgs_",
                "nonlinearoptimization_mp_conjugategradient_", "nonlinearoptimization_mp_conjugategradient_basic_"):
        assert sym in exported


def test_library_loads_without_gpu_and_reports_version():
    _build()
    import fortran_library_b200 as fl
    L = fl.lib()
    assert b"sm_100a" in L.flgpu_version()
    assert fl.device_count() >= 0


def test_sass_is_sm_100a_fp64():
    """The shipped cubins are sm_100a and the streaming kernels use 128-bit loads and DFMA."""
    _build()
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    dot = sass[sass.index("dot_kernel"):]
    dot = dot[:dot.index("Function :", 10) if "Function :" in dot[10:] else len(dot)]
    assert "DFMA" in dot and "LDG.E.128" in dot
    # K3 streams its columns with 1-D bulk async copies (cp.async.bulk -> UBLKCP) signalled on mbarriers
    assert "UBLKCP" in sass and "SYNCS.ARRIVE.TRANS64" in sass


def test_struct_layout_matches_ctypes(tmp_path):
    import fortran_library_b200._capi as capi
    prog = tmp_path / "layout.c"
    fields = {
        "flgpu_options": [f for f, _ in capi.Options._fields_],
        "flgpu_stats": [f for f, _ in capi.Stats._fields_],
        "flgpu_iter_info": [f for f, _ in capi.IterInfo._fields_],
        "flgpu_eval_ctx": [f for f, _ in capi.EvalCtx._fields_],
        "flgpu_problem": [f for f, _ in capi.Problem._fields_],
        "flgpu_constraints": [f for f, _ in capi.Constraints._fields_],
        "flgpu_al_options": [f for f, _ in capi.ALOptions._fields_],
        "flgpu_al_stats": [f for f, _ in capi.ALStats._fields_],
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for s, fs in fields.items():
        lines.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for f in fs:
            lines.append(f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    lines.append("return 0;}")
    prog.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", str(prog), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    mirror = {"flgpu_options": capi.Options, "flgpu_stats": capi.Stats, "flgpu_iter_info": capi.IterInfo,
              "flgpu_eval_ctx": capi.EvalCtx, "flgpu_problem": capi.Problem, "flgpu_constraints": capi.Constraints,
              "flgpu_al_options": capi.ALOptions, "flgpu_al_stats": capi.ALStats}
    for s, cls in mirror.items():
        assert int(got[s]) == C.sizeof(cls), s
        for f, _ in cls._fields_:
            assert int(got[f"{s}.{f}"]) == getattr(cls, f).offset, f"{s}.{f}"


def test_cpp_wrappers_link(tmp_path):
    """tests/link/cpp_dropin.cpp (reference test.cpp style, through include/NonlinearOptimization_flgpu.hpp)
    links against libflgpu.so leaving no undefined optimizer symbol."""
    _build()
    exe = tmp_path / "cpp_dropin"
    subprocess.run(["g++", "-std=c++11", os.path.join(ROOT, "tests", "link", "cpp_dropin.cpp"), "-o", str(exe),
                    "-L" + os.path.dirname(LIB), "-lflgpu", "-Wl,-rpath," + os.path.dirname(LIB)], check=True)
    undefined = subprocess.run(["nm", "-u", str(exe)], check=True, capture_output=True, text=True).stdout
    assert "__nonlinearoptimization_MOD_lbfgs" in undefined  # resolved at load time from libflgpu.so
    subprocess.run(["ldd", str(exe)], check=True, capture_output=True)


@pytest.mark.skipif(not os.path.exists("/root/reference/cpp/NonlinearOptimization.hpp"),
                    reason="the reference tree is only mounted in the build container")
def test_unmodified_reference_header_links(tmp_path):
    """A program using the UNMODIFIED reference header's SteepestDescent, ConjugateGradient and AugmentedLagrangian
    wrappers (hpp:395-454, 514-545) links against libflgpu.so: the binary drop-in claim for this path."""
    if not os.path.exists("/root/reference/cpp/NonlinearOptimization.hpp"):
        pytest.skip("the reference tree is not mounted here")
    _build()
    src = tmp_path / "ref_hdr.cpp"
    src.write_text(r'''
#include <string>
#include <tuple>
#include "/root/reference/cpp/NonlinearOptimization.hpp"
static void f(double & fx, const double * x, const int & dim) { fx = 0; for (int i = 0; i < dim; i++) fx += x[i]*x[i]*x[i]*x[i]; }
static void fd(double * g, const double * x, const int & dim) { for (int i = 0; i < dim; i++) g[i] = 4*x[i]*x[i]*x[i]; }
static int f_fd(double & fx, double * g, const double * x, const int & dim) { f(fx, x, dim); fd(g, x, dim); return 0; }
static void c(double * cx, const double * x, const int & M, const int & N) { cx[0] = -1; for (int i = 0; i < N; i++) cx[0] += x[i]*x[i]; (void)M; }
static void cd(double * cdx, const double * x, const int & M, const int & N) { for (int i = 0; i < N; i++) cdx[i] = 2*x[i]; (void)M; }
int main() { double x[10] = {0.5}; FL::NO::ConjugateGradient(f, fd, x, 10); FL::NO::ConjugateGradient(f, fd, f_fd, x, 10, "PR");
  FL::NO::SteepestDescent(f, fd, f_fd, x, 10);
  FL::NO::AugmentedLagrangian(f, fd, f_fd, nullptr, c, cd, nullptr, x, 10, 1, "LBFGS"); return 0; }
''')
    exe = tmp_path / "ref_hdr"
    subprocess.run(["g++", "-std=c++11", str(src), "-o", str(exe), "-L" + os.path.dirname(LIB), "-lflgpu",
                    "-Wl,-rpath," + os.path.dirname(LIB)], check=True)


def test_augmented_lagrangian_forwards_dense_solvers_to_libfl(tmp_path):
    """UnconstrainedSolver = 'BFGS' / 'NewtonRaphson' are outside the GPU path: libflgpu hands the call to the next
    definition of __nonlinearoptimization_MOD_augmentedlagrangian (libFL linked after it).  Needs no GPU."""
    fake = tmp_path / "fakefl.c"
    fake.write_text(
        '#include <stdio.h>\n'
        'void __nonlinearoptimization_MOD_augmentedlagrangian(void *f, void *fd, void *c, void *cd, double *x, const int *N,\n'
        '    const int *M, const char *solver, const double *l0, const double *m0, void *fdd, void *cdd, const int *es,\n'
        '    const int *mem, const char *meth, void *ffd, const int *s, const int *w, const int *mi, const double *p,\n'
        '    const double *ms, const double *c1, const double *c2, const double *inc, int ls, int lm) {\n'
        '    printf("libFL got %.*s N=%d M=%d\\n", ls, solver, *N, *M); x[0] = 42.0; }\n')
    prog = tmp_path / "prog.cpp"
    prog.write_text(
        '#include <cstdio>\n#include "%s"\n'
        'static void f(double &fx, const double *, const int &) { fx = 0; }\n'
        'static void fd(double *, const double *, const int &) {}\n'
        'static void c(double *, const double *, const int &, const int &) {}\n'
        'int main() { double x[3] = {1, 2, 3};\n'
        '  FL::NO::AugmentedLagrangian(f, fd, nullptr, nullptr, c, c, nullptr, x, 3, 1, "BFGS");\n'
        '  std::printf("x0=%%g\\n", x[0]); return x[0] == 42.0 ? 0 : 1; }\n'
        % os.path.join(ROOT, "include", "NonlinearOptimization_flgpu.hpp"))
    libdir = os.path.join(ROOT, "fortran_library_b200")
    subprocess.run(["gcc", "-shared", "-fPIC", str(fake), "-o", str(tmp_path / "libfakeFL.so")], check=True)
    exe = tmp_path / "prog"
    # --no-as-needed: this toy program needs nothing else from "libFL", a real one does (every other module)
    subprocess.run(["g++", "-std=c++11", str(prog), "-o", str(exe), "-Wl,--no-as-needed", "-L" + libdir, "-lflgpu",
                    "-L" + str(tmp_path), "-lfakeFL", "-Wl,-rpath," + libdir, "-Wl,-rpath," + str(tmp_path)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "libFL got BFGS N=3 M=1" in r.stdout, r.stdout + r.stderr
    # without a libFL behind it: the message, exit status 1
    exe2 = tmp_path / "prog2"
    subprocess.run(["g++", "-std=c++11", str(prog), "-o", str(exe2), "-L" + libdir, "-lflgpu", "-Wl,-rpath," + libdir],
                   check=True)
    r = subprocess.run([str(exe2)], capture_output=True, text=True)
    assert r.returncode == 1 and "dense-Hessian" in r.stdout


def test_objective_helper_header_compiles_for_sm_100a(tmp_path):
    """include/flgpu_objective.cuh (+ flgpu_search_core.hpp) cross-compiles with nvcc for sm_100a without a GPU: the
    user-objective fixture the GPU tests run, its four callbacks and the cooperative search kernel."""
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    _build()
    out = tmp_path / "libuser_objective.so"
    libdir = os.path.join(ROOT, "fortran_library_b200")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-std=c++17", "-Xcompiler",
                    "-fPIC", "-shared", "-o", str(out), os.path.join(ROOT, "tests", "link", "user_objective.cu"),
                    "-L" + libdir, "-lflgpu", "-Xlinker", "-rpath," + libdir], check=True)
    sass = subprocess.run(["cuobjdump", "-sass", str(out)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass and "search_kernel" in sass and "DADD" in sass
    lib = C.CDLL(str(out))                       # loads (against libflgpu.so) and exports the fixture's entry point
    assert hasattr(lib, "user_problem")


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point aborts with a message; nothing is computed on the
    CPU.  (Skipped on a GPU box, where the same call simply runs.)"""
    _build()
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import numpy as np, fortran_library_b200 as fl\n"
            "if fl.device_count() > 0: sys.exit(77)\n"
            "import ctypes as C\n"
            "p = fl.capi.Problem(); fl.lib().flgpu_builtin_problem(0, C.byref(p))\n"
            "o = fl.default_options(); st = fl.capi.Stats(); x = np.ones(4)\n"
            "fl.lib().flgpu_lbfgs(C.byref(p), C.byref(o), x.ctypes.data, 4, 0, C.byref(st))\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    if r.returncode == 77:
        pytest.skip("a GPU is present")
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr
    # and the Python layer refuses before reaching C
    import fortran_library_b200 as fl
    if fl.device_count() == 0:
        with pytest.raises(fl.FlgpuError):
            fl.LBFGS(fl.capi.Problem(), __import__("numpy").ones(4))


def test_product_never_references_the_oracle():
    """The product tree must not import, link or load anything under oracle/ or tests/hostsim/."""
    pkg = os.path.join(ROOT, "fortran_library_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "liboracle" not in text and "oracle/" not in text.replace("the oracle", ""), fn
                assert "hostsim" not in text or fn in ("backend.hpp",), fn
    needed = subprocess.run(["readelf", "-d", LIB], capture_output=True, text=True).stdout
    assert "oracle" not in needed and "hostsim" not in needed
