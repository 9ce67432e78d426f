"""BASELINE configs[0] wall clock: the reference executable (host cores) and ours (GPU) on the solar-system CSV,
dt = 1h, vs = 1d, naive opt_stage 2 / Barnes-Hut, for several t_end (fixed start-up cost vs per-step cost).
python tools/config1_timing.py [t_end ...]"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fixture = os.path.join(ROOT, "tests", "golden", "solar_178.csv")
exes = {"reference (oracle/_ref, host cores)": os.path.join(ROOT, "oracle", "_ref", "N_Body_Simulation"),
        "ours (libnbody_b200.so)": os.path.join(ROOT, "n-body-simulation_b200", "N_Body_Simulation"),
        "reference driver + our back end (integration/)": os.path.join(ROOT, "integration", "_build", "N_Body_Simulation_b200")}
ends = sys.argv[1:] or ["365d"]
for alg in ("naive", "BarnesHut"):
    for name, exe in exes.items():
        for t_end in ends:
            with tempfile.TemporaryDirectory() as d:
                t0 = time.perf_counter()
                r = subprocess.run([exe, "--file=" + fixture, "--dt=1h", "--t_end=" + t_end, "--vs=1d", "--vs_dir=" + d,
                                    "--algorithm=" + alg, "--opt_stage=2"], capture_output=True, text=True)
                dt = time.perf_counter() - t0
                print("%-9s %-48s t_end=%-5s rc=%d  %.2f s wall" % (alg, name, t_end, r.returncode, dt), flush=True)
