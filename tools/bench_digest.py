"""Human-readable digest of a bench.py JSON line: tools/bench_digest.py gpurun_out/bench.json"""
import json, sys
txt = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]
d = json.loads(txt)
print("headline: value %.0f %s, %.3f ms/step, e2e %.0f (%.3f ms), launches %s, kernel %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["config"].get("kernel")))
r = d["roofline"]
print("roofline: achieved %.1f GB/s of %.1f -> frac %.4f; onchip frac %.3f; kernel_ms %.3f; traffic %s" % (r["achieved"], r["peak"], r["frac"], r["onchip"]["frac_of_128B_per_clk_per_SM"], r["kernel_ms"], r["traffic"]))
print("clocks:", d["clocks"])
print("cpu:", d.get("cpu_baseline"))
for row in d.get("sweep_rows", []):
    print("ROW %-62s %-5s %-60s %8.0f/s %8.3f ms e2e %8.0f hbm %.4f onchip %.3f zero %d fb %d parity %s" % (
        row["workload"][:62], row["kernel"], json.dumps(row["plan"]) if row["plan"] else "-", row["value"], row["ms_per_step"], row["e2e"]["value"],
        row["roofline"]["frac"], row["roofline"]["onchip_frac_of_128B_per_clk_per_SM"], row["zero_volume_pairs"], row["single_match_fallbacks"], row["parity_exact"]))
if "cfg5" in d:
    for row in d["cfg5"]["rows"]:
        print("CFG5", row)
for k, v in (d.get("seq_match") or {}).items():
    print("SEQ", k, v)
if "replay" in d:
    print("REPLAY", json.dumps(d["replay"])[:1200])
g = d.get("graph_solve")
if g:
    for tag, x in (("0.05/0.02", g), ("0.03/0.01", g.get("low_noise_variant", {}))):
        if not x:
            continue
        print("GRAPH", tag, "ms %.2f wall %.2f setup %.3f lm %d/%d pcg %d cost %.3f" % (x["ms"], x["wall_ms"], x["host_setup_ms"], x["lm_iterations"], x["successful_steps"], x["pcg_iterations"], x["final_cost"]))
        print("      roofline frac %.4f, us/pcg-iter %.2f; incremental %s" % (x["roofline"]["frac"], x["roofline"]["us_per_pcg_iteration_incl_everything"], x["incremental"]))
        print("      cpu %s parity %s" % ({k: x.get("cpu_baseline", {}).get(k) for k in ("ms", "kind", "lm_iterations")}, x.get("parity_vs_oracle")))
print("OG", d.get("occupancy_grid"))
