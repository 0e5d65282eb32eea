"""Host logic of the network compiler (bnbp_jit.cu): runs WITHOUT a GPU -- source generation and
the NVRTC compile to an sm_100a cubin are pure host work; no kernel is launched here."""
import os
import re
import subprocess

import pytest

from bayesiannetwork_b200 import synth


@pytest.fixture(scope="module")
def engine():
    from bayesiannetwork_b200 import _build, engine
    _build.build()
    return engine


def test_generated_source_describes_the_network(engine):
    net = synth.pearl_network()
    src = engine.spec_source(net, "fp64", 2)
    assert "#define BNBP_T double" in src and "#define BNBP_VARIANT 2" in src
    assert "#define BNBP_NCPT 16" in src
    assert len(re.findall(r"^struct N\d+ ", src, flags=re.M)) == 4
    # H has parents R, S (cards 2, 2) and no children (libs/bayesian/test/belief_propagation.cpp:9-62)
    n3 = [l for l in src.splitlines() if l.startswith("struct N3 ")][0]
    assert "R=2,K=2,M=0" in n3 and "RU[2]={2,2}" in n3
    # software pipeline: every node is declared, loaded and computed exactly once, loads first
    walk = [l for l in src.splitlines() if l.startswith("#define BNBP_WALK")][0]
    for i in range(4):
        assert walk.count(f"BNBP_DECL(N{i})") == 1 and walk.count(f"BNBP_COMP(N{i})") == 1
        assert walk.index(f"BNBP_LOAD(N{i})") < walk.index(f"BNBP_COMP(N{i})")
    assert "extern \"C\" __global__" in src                      # the hand-written body follows
    assert engine.spec_source(net, "fp32", 0).count("#define BNBP_T float") == 1


def test_precompile_pearl_into_cache(engine, tmp_path, monkeypatch):
    """NVRTC -> cubin -> cache file, no GPU involved; the cubin is sm_100a and holds the kernel."""
    monkeypatch.setenv("BNBP_CACHE_DIR", str(tmp_path))
    engine.precompile(synth.pearl_network(), "fp64", 0b001)
    files = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert len(files) == 1
    engine.precompile(synth.pearl_network(), "fp64", 0b001)      # served by the cache
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".cubin")]) == 1
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if os.path.exists(cuobjdump):
        out = subprocess.run([cuobjdump, "-elf", str(tmp_path / files[0])], capture_output=True, text=True).stdout
        assert "bnbp_spec_sweep" in out and "bnbp_cpt" in out
        sass = subprocess.run([cuobjdump, "-sass", str(tmp_path / files[0])], capture_output=True, text=True).stdout
        assert "sm_100" in sass
        assert "c[0x3]" in sass                                   # CPT entries are constant-bank operands


def test_ineligible_networks_are_reported(engine):
    from bayesiannetwork_b200.engine import BnbpError
    with pytest.raises(BnbpError, match="not eligible"):
        engine.precompile(synth.high_card(6, card=32, n_parents=2, seed=10), "fp64", 1)
    # 1600 nodes: too many to unroll node by node -- walked class by class instead (round 2, BNBP_CLASSLOOP) ...
    assert "#define BNBP_CLASSLOOP 1" in engine.spec_source(synth.grid(40), "fp64", 0)
    # ... unless that is switched off, or the node shapes are too many / too large for it as well
    os.environ["BNBP_CLASSLOOP"] = "0"
    try:
        with pytest.raises(BnbpError, match="not eligible.*1024 nodes"):
            engine.spec_source(synth.grid(40), "fp64", 0)
    finally:
        del os.environ["BNBP_CLASSLOOP"]
    with pytest.raises(BnbpError, match="not eligible.*class-looped walk"):
        engine.spec_source(synth.random_dag(1500, max_parents=4, card_lo=2, card_hi=8, seed=3), "fp64", 0)


def test_cmake_project_configures(tmp_path):
    """CMakeLists.txt (the CUDA / C-ABI build wiring the north star asks for) configures with the
    image's cmake + nvcc; the full build is exercised by hand (takes a minute), not here."""
    import shutil
    cmake = shutil.which("cmake")
    if not cmake or not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("cmake / nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    r = subprocess.run([cmake, "-S", root, "-B", str(tmp_path / "b"), "-DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc"],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert (tmp_path / "b" / "gen" / "bnbp_spec_embed.inc").exists()


def test_first_use_compile_cost_decides_between_the_generators(engine, tmp_path, monkeypatch):
    """NVRTC time of an unrolled walk grows like N^2.4 (100 nodes: 144 s per variant), so above 64 nodes the unrolled
    kernels are taken only from the cubin cache; a network that can be walked class by class gets that generator instead
    (100-node grid: 6 classes, one second)."""
    monkeypatch.setenv("BNBP_CACHE_DIR", str(tmp_path))           # an empty cache
    small, mid = synth.grid(8), synth.grid(10)                    # 64 and 100 nodes, both within the unrolled walk's own bounds
    assert "#define BNBP_CLASSLOOP 1" not in engine.spec_source(small, "fp64", 0)
    src = engine.spec_source(mid, "fp64", 0)
    walk = [l for l in src.splitlines() if l.startswith("#define BNBP_WALK")][0]
    assert "#define BNBP_CLASSLOOP 1" in src and 4 <= walk.count("BNBP_CLASS(") <= 9
    monkeypatch.setenv("BNBP_CLASSLOOP", "0")                     # never class-looped: the unrolled source is still on offer
    assert "struct N99 " in engine.spec_source(mid, "fp64", 0)    # (AUTO runs keep the generic kernel until it is compiled once)
