"""Compile the REFERENCE's own Chamfer / EMD CUDA extensions (metrics/chamfer_dist/{chamfer.cu,chamfer_cuda.cpp},
metrics/PyTorchEMD/cuda/{emd_kernel.cu,emd.cpp}) from where they lie under /root/reference into oracle/_ref/ -- TEST
INFRASTRUCTURE (see oracle/__init__.py): tests/test_gpu_metrics.py uses them on the GPU box as the checker that pins
row N4 (csrc/metrics.cu) against the reference itself.  No reference source is copied; only the built modules
(git-ignored, they travel with the gpurun snapshot) land in oracle/_ref/.  Run where /root/reference exists:

    python -m oracle.build_ref_metrics          (also called by __graft_entry__.build())
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
MODULES = {
    "ref_chamfer": ("metrics/chamfer_dist", ["chamfer.cu", "chamfer_cuda.cpp"]),
    "ref_emd_cuda": ("metrics/PyTorchEMD/cuda", ["emd_kernel.cu", "emd.cpp"]),
}


def build(ref_root: str = "/root/reference", verbose: bool = False) -> list:
    if not os.path.isdir(os.path.join(ref_root, "metrics")):
        return []
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    inc = list(ce.include_paths("cuda")) + [sysconfig.get_paths()["include"]]
    incs = sum((["-I", i] for i in inc), [])
    abi = "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    built = []

    def one(name_sub_files):
        name, (sub, files) = name_sub_files
        so = os.path.join(OUT, name + sysconfig.get_config_var("EXT_SUFFIX"))
        srcs = [os.path.join(ref_root, sub, f) for f in files]
        if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
            return so
        objs = []
        for s in srcs:
            o = os.path.join(OUT, name + "_" + os.path.basename(s) + ".o")
            common = ["-std=c++17", f"-DTORCH_EXTENSION_NAME={name}", "-DTORCH_API_INCLUDE_EXTENSION_H", abi, *incs]
            if s.endswith(".cu"):
                cmd = [nvcc, "-c", s, "-o", o, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-Xcompiler", "-fPIC", *common]
            else:
                cmd = ["g++", "-c", s, "-o", o, "-O2", "-fPIC", *common]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"building the reference's {os.path.basename(s)} failed:\n{r.stderr[-2000:]}")
            objs.append(o)
        link = ["g++", "-shared", "-o", so, *objs, f"-L{tlib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
                "-ltorch_python", "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{tlib}"]
        r = subprocess.run(link, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"linking {name} failed:\n{r.stderr[-2000:]}")
        for o in objs:
            os.remove(o)
        if verbose:
            print("built", so)
        return so

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(MODULES)) as ex:          # the two .cu files take ~3 minutes each (torch headers)
        built = list(ex.map(one, MODULES.items()))
    return built


def load(name: str):
    """import oracle/_ref/<name>.so (None when it was not built)"""
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    so = os.path.join(OUT, name + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose=True))
