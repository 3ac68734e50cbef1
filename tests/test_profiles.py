"""CPU tests of the evidence plumbing: the committed ncu launch list, the tool that derives the per-step DRAM traffic from it,
and the hash that lets bench.py quote that figure only for the kernel sources it was measured on."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_step_traffic_file_matches_the_committed_launch_list():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_step_traffic.py"), os.path.join(ROOT, "profiles", "r2_launches.csv")],
                         check=True, capture_output=True, text=True).stdout
    derived = json.loads(out)
    with open(os.path.join(ROOT, "profiles", "r2_step_dram_traffic.json")) as f:
        committed = json.load(f)
    for k in ("conv_tc_launches", "conv_tc_dram_bytes_per_step", "pool_conv_tc_launches", "pool_conv_tc_dram_bytes_per_step", "step_time"):
        assert derived[k] == committed[k], k
    # one step = input transform, 21 conv launches, 3 MaxPool-fused squeeze launches, GAP + Softmax
    assert committed["conv_tc_launches"] == 21 and committed["pool_conv_tc_launches"] == 3
    assert len(committed["launches"]) == 26
    # DRAM traffic of the conv launches stays below their algorithmic bytes (L2 reuse between consecutive launches), and the
    # pool-fused launches read the pool inputs once (1.17 + 0.76 + 0.38 GB) and write the squeeze outputs
    assert 4.5e9 < committed["conv_tc_dram_bytes_per_step"] < 7.3e9
    assert 2.3e9 < committed["pool_conv_tc_dram_bytes_per_step"] < 2.6e9


def test_committed_traffic_figure_is_for_the_current_kernel_sources():
    """bench.py prints roofline.traffic = null when this does not hold; here it is a skip with the recipe, not a failure of the code."""
    import pytest
    import bench
    with open(os.path.join(ROOT, "profiles", "r2_step_dram_traffic.json")) as f:
        committed = json.load(f)
    if bench.kernel_sources_sha() != committed["kernel_sources_sha"]:
        pytest.skip("kernel code changed since profiles/r2_step_dram_traffic.json was taken: re-run tools/exp/artifacts_r2.sh "
                    "under gpurun and tools/collect_r2_profiles.sh")


def test_kernel_source_hash_ignores_comments_only(tmp_path, monkeypatch):
    import bench
    base = bench.kernel_sources_sha()
    committed = {"kernel_sources_sha": base}
    # same code with extra comments / blank lines -> same hash; a changed token -> different hash
    src = os.path.join(ROOT, "onnx_rusty_inference_engine_b200", "csrc")
    fake = tmp_path / "onnx_rusty_inference_engine_b200" / "csrc"
    fake.mkdir(parents=True)
    for name in bench.KERNEL_SOURCES:
        text = open(os.path.join(src, name), encoding="utf-8", errors="replace").read()
        (fake / name).write_text("// a comment\n\n" + text + "\n/* another\n one */\n", encoding="utf-8")
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.kernel_sources_sha() == committed["kernel_sources_sha"]
    first = fake / bench.KERNEL_SOURCES[0]
    first.write_text(first.read_text(encoding="utf-8").replace("constexpr int BM = 128;", "constexpr int BM = 64;", 1), encoding="utf-8")
    assert bench.kernel_sources_sha() != committed["kernel_sources_sha"]
