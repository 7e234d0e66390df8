"""CPU checks of what bench.py promises the driver: both arms describe the workload with the same `config`, and every traffic figure the
line can carry names a capture that is committed under profiles/."""
import importlib.util
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_both_arms_print_the_same_config_keys():
    bench = _bench()
    src = open(os.path.join(ROOT, "bench.py")).read()
    # the two JSON lines of the default mode (ours, reference) take their "config" from the same helper (the batch and row-band modes add their own keys)
    assert len(re.findall(r'"config": common_config\(args\.workload\)', src)) == 2
    for workload in bench.WORKLOADS:
        cfg = bench.common_config(workload)
        assert set(cfg) == {"workload", "width", "height", "channels", "error_factor", "fast_bit_crushing"}
        assert cfg["workload"] == workload and cfg["channels"] in (3, 4)


def test_traffic_figures_name_committed_captures():
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
        table = json.load(f)
    assert table, "no capture recorded"
    for workload, phases in table.items():
        for phase, entry in phases.items():
            assert int(entry["bytes"]) > 0, (workload, phase)
            m = re.match(r"(profiles/[\w.\-]+)", entry["source"])
            assert m, entry["source"]
            assert os.path.isfile(os.path.join(ROOT, m.group(1))), m.group(1)


def test_ncu_traffic_lookup_returns_none_without_a_capture():
    bench = _bench()
    assert bench.ncu_traffic("no_such_workload", "merge_scan") is None
    got = bench.ncu_traffic("c2_4k_photo", "merge_scan")
    assert got is not None and got["bytes"] > 0
