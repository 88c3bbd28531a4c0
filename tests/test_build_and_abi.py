"""CPU-side checks: constants file is what the oracle derives, the device limb schedule passes its host
unit test, and libb200zk.so exports every symbol include/b200zk.h declares (no compute without a GPU)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_constants_inc_is_derived_from_oracle():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_constants", os.path.join(ROOT, "tools", "gen_constants.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    path = os.path.join(ROOT, "distributed_groth16_b200", "csrc", "bn254_constants.inc")
    assert open(path).read() == mod.render()


def test_device_limb_schedule_on_host(cref, tmp_path):
    """csrc/fp.cuh + ec.cuh compiled for the host (PTX carry primitives emulated) vs the oracle."""
    exe = tmp_path / "fp_host_test"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-frounding-math", "-o", str(exe), os.path.join(ROOT, "tests", "host", "fp_host_test.cpp"),
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
                           "-fopenmp"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout + out.stderr


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200zk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200zk_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from distributed_groth16_b200 import _native, build
    build.build()
    lib = _native.lib()                      # sets argtypes for every entry of SIGNATURES (AttributeError if missing)
    declared = _declared_symbols()
    assert len(declared) >= 25
    assert sorted(_native.SIGNATURES) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.b200zk_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from distributed_groth16_b200 import B200zkError, Net
    with pytest.raises(B200zkError):
        Net(0)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may touch oracle/: an AST scan of every product
    module for imports of `oracle`, and a text scan of the native sources for includes of it."""
    import ast
    pkg = os.path.join(ROOT, "distributed_groth16_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                for node in ast.walk(ast.parse(open(path).read(), path)):
                    mods = []
                    if isinstance(node, ast.Import):
                        mods = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        mods = [node.module or ""]
                    offenders += [(path, m) for m in mods if m == "oracle" or m.startswith("oracle.")]
            elif f.endswith((".cu", ".cuh", ".h", ".inc")):
                offenders += [(path, line.strip()) for line in open(path).read().splitlines()
                              if line.lstrip().startswith("#include") and "oracle" in line]
    assert not offenders, offenders
    # bench.py: oracle imports only inside run_reference / the cpu_baseline legs (functions that say so), never at module level
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not [n for n in top if "oracle" in ast.dump(n)]
