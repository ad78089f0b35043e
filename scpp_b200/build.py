"""Builds scpp_b200/libscpp_b200.so (CUDA, sm_100a) in-tree with nvcc.  No JIT cache: the .so travels with the repo snapshot.
The heavy kernels are explicit instantiations compiled as separate translation units in parallel (csrc/kernels_inst.cu)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
OUT = os.path.join(HERE, "libscpp_b200.so")
# -static-global-template-stub=false: the kernels are explicit instantiations DEFINED in other translation units (kernels_inst.cu) and only
# declared (extern template) where they are launched
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-static-global-template-stub=false"]
GROUPS = [(m, g) for m in (0, 3, 1, 2) for g in range(6)]      # models: RocketQuat, RocketQuatRollPlugin (the slow ones first), Rocket2d, Rocket2dPlugin


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(HERE, "..", "include", f) for f in ("scpp_b200.h", "scpp_cvx.hpp", "scpp_plugin.hpp")]
    deps += [os.path.join(HERE, "plugins", f) for f in os.listdir(os.path.join(HERE, "plugins"))] + [os.path.join(HERE, "..", "tools", "gen_plugin.cpp")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def generate_plugins():
    """models written against the plugin surface (scpp_b200/plugins/): record addApplicationConstraints through the cvx:: shim and write the
    stage-wise tables csrc/gen/*.inc (tools/gen_plugin.cpp); runs before nvcc, rewrites a file only when its content changes"""
    root = os.path.dirname(HERE)
    exe = os.path.join(HERE, "_obj", "gen_plugin")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    os.makedirs(os.path.join(SRC, "gen"), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"), os.path.join(root, "tools", "gen_plugin.cpp"), "-o", exe])
    subprocess.check_call([exe, os.path.join(SRC, "gen")])


def build(force=False, verbose=False, variant=None, defines=(), only=None):
    """variant: A/B experiments -- builds libscpp_b200_<variant>.so with extra -D defines (selected at run time with SCPP_B200_LIB)"""
    global OBJ, OUT
    if variant:
        OBJ = os.path.join(HERE, "_obj_" + variant); OUT = os.path.join(HERE, f"libscpp_b200_{variant}.so"); force = True
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    generate_plugins()
    extra = (["-Xptxas", "-v"] if verbose else []) + list(defines)
    jobs = [([nvcc] + ARCH + extra + ["-c", os.path.join(SRC, "engine.cu"), "-o", os.path.join(OBJ, "engine.o")], "engine"),
            ([nvcc] + ARCH + extra + ["-c", os.path.join(SRC, "mpc.cu"), "-o", os.path.join(OBJ, "mpc.o")], "mpc")]
    if variant and only is not None:      # rebuild only the listed (model, group) kernel objects; the rest comes from the main build
        import shutil
        main_obj = os.path.join(HERE, "_obj")
        for f in os.listdir(main_obj):
            shutil.copy2(os.path.join(main_obj, f), os.path.join(OBJ, f))
        jobs = []
    for m, g in GROUPS:
        if variant and only is not None and (m, g) not in only:
            continue
        jobs.append(([nvcc] + ARCH + extra + [f"-DSCPP_KERNEL_MODEL={m}", f"-DSCPP_KERNEL_GROUP={g}", "-c", os.path.join(SRC, "kernels_inst.cu"),
                                              "-o", os.path.join(OBJ, f"k_{m}_{g}.o")], f"kernels model {m} group {g}"))

    def run(job):
        cmd, name = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(f"== {name}\n{r.stdout}{r.stderr}")
        if r.returncode:
            raise RuntimeError(f"nvcc failed: {name}")

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, "engine.o"), os.path.join(OBJ, "mpc.o")] + [os.path.join(OBJ, f"k_{m}_{g}.o") for m, g in GROUPS]
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs + ["-ldl"])
    return OUT


if __name__ == "__main__":
    if "--variant" in sys.argv:      # python scpp_b200/build.py --variant w8 -DSCPP_WPB_MAX=8
        i = sys.argv.index("--variant")
        only = [(0, 0)] if "--k2-only" in sys.argv else None      # only k_solve<RocketQuat> (kernel group 0 of model 0)
        print(build(variant=sys.argv[i + 1], defines=[a for a in sys.argv[i + 2:] if a.startswith("-D")] + [f for a in sys.argv[i + 2:] if a.startswith("--nvcc=") for f in a[7:].split(",")], verbose="-v" in sys.argv, only=only))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
