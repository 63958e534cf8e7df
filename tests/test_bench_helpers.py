"""bench.py host logic that runs without a GPU: the JSON config block, the clock-sample summary, the scenario
chunking used for the synthetic CDU workload."""
import argparse

import numpy as np

import bench


def test_clock_sampler_summary_parses_nvidia_smi_rows():
    cs = bench.ClockSampler(0)
    cs.rows = [["1965", "1965", "700.1", "Not Active", "Not Active", "Not Active", "Active"],
               ["1700", "1965", "990.0", "Not Active", "Not Active", "Not Active", "Active"],
               ["[N/A]", "1965", "1", "Not Active", "Not Active", "Not Active", "Not Active"],
               ["1800", "1965", "800.0", "Not Active", "Active", "Not Active", "Not Active"]]
    s = cs.summary()
    assert s["sm_mhz"] == 1800.0 and s["sm_max_mhz"] == 1965.0 and s["samples"] == 3
    assert s["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]
    assert bench.ClockSampler(0).summary()["samples"] == 0


def test_config_block_names_the_workload():
    args = argparse.Namespace(horizon=140, gpus=2, precision="mixed", slots=16384, unique_slabs=6)
    c = bench._config(args, 65536, 16)
    assert c["qp_vars"] == 4480 and c["trajectories_per_gpu"] == 65536 and c["concurrent_slots_per_gpu"] == 16384
    assert "configs[2]" in c["workload"] and "model" not in c and c["sim_steps_per_step"] == 16


def test_scenarios_are_contiguous_chunks_of_one_prbs_signal():
    p, sp, ds = bench._scenarios(5, 400, seed=3)
    assert sp.shape == (5, 400, p.Ny) and ds.shape == (5, 400, p.Nd)
    assert np.array_equal(sp.reshape(2000, p.Ny), p.setpoints) and np.array_equal(ds.reshape(2000, p.Nd), p.disturbances)
    # piecewise constant with the reference's hold statistics (mean 400 / 200 steps): few distinct rows
    assert len(np.unique(p.setpoints, axis=0)) < 20 and len(np.unique(p.disturbances, axis=0)) < 40
