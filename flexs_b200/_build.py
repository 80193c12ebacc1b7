"""Builds ``flexs_b200/libflexs_b200.so`` (sm_100a only) with nvcc, in-tree.

The library is a plain C-ABI shared object (include/flexs_b200.h); it links the CUDA runtime
statically and nothing else, so it can be loaded next to torch or from any FFI.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path
from typing import Optional

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libflexs_b200.so"

SOURCES = ["api.cu", "encode.cu", "cnn_simple.cu", "cnn_tiled.cu", "cnn_umma.cu", "cnn_umma2.cu", "cnn_k9.cu", "cnn_a20.cu", "mlp.cu", "mlp_umma.cu",
           "topk.cu", "select.cu", "peer.cu", "dedup.cu", "gen.cu", "density.cu", "train.cu", "vae.cu", "landscape.cu", "enum_table.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    lib_m = LIB_PATH.stat().st_mtime
    deps = (list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) +
            [PKG_DIR.parent / "include" / "flexs_b200.h"])
    return any(p.stat().st_mtime > lib_m for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library."""
    if not force and not needs_build():
        build_packstr()
        return LIB_PATH
    obj_dir = CSRC / "build"
    obj_dir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = obj_dir / (src[:-3] + ".o")
        objs.append(str(obj))
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, cmd, proc in procs:
        out, _ = proc.communicate()
        log.append(f"$ {' '.join(cmd)}\n{out}")
        if proc.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
    (obj_dir / "ptxas.log").write_text("\n".join(log))
    if failed:
        raise RuntimeError("flexs_b200: nvcc compilation failed (see stderr)")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
            "-cudart", "static", "-o", str(LIB_PATH), *objs]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("flexs_b200: link failed")
    if verbose:
        print("\n".join(log))
    build_packstr(force)
    return LIB_PATH


def packstr_path() -> Path:
    import sysconfig

    return PKG_DIR / ("_packstr" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_packstr(force: bool = False) -> Optional[Path]:
    """Compile the CPython host helper (csrc/packstr.c, no CUDA) with the C compiler, in-tree.

    The helper is an accelerator of the host boundary only (strings -> byte matrix / packed residues in one C pass);
    ``sequence_utils`` has an equivalent numpy route.  A box without a C compiler or Python.h therefore gets a
    warning and ``None``, never a failed ``build()``."""
    import sysconfig
    import warnings

    out = packstr_path()
    src = CSRC / "packstr.c"
    if not force and out.exists() and out.stat().st_mtime >= src.stat().st_mtime:
        return out
    cc = os.environ.get("CC") or sysconfig.get_config_var("CC") or "gcc"
    cmd = [*cc.split(), "-O3", "-shared", "-fPIC", "-pthread", f"-I{sysconfig.get_paths()['include']}", str(src), "-o", str(out)]
    try:
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    except OSError as e:
        warnings.warn(f"flexs_b200: host helper _packstr not built ({e}); using the numpy route")
        return None
    if res.returncode != 0:
        warnings.warn(f"flexs_b200: host helper _packstr not built; using the numpy route\n{res.stdout}")
        return None
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
