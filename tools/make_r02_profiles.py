"""Rebuilds profiles/r02_ncu_summary.md and profiles/roofline_traffic.json from the round-2 captures under gpurun_out/
(r02_ncu_fast_c3, r02_ncu_encode_c2b, r02_ncu_tile_c256, r02_ncu_iwt: ncu --set full; r02_launches_bench.csv: launch list) and
copies the bench lines they belong to.  Development tool, runs without a GPU."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max"]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    return {h[i]: (v[i], u[i]) for i in range(len(h))}


def bytes_of(d, k):
    val, unit = d[k]
    return float(val) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit]


def section(title, rep, what, alg_bytes=None):
    d = load(rep)
    tab = "| metric | value | unit |\n|---|---|---|\n" + "".join(f"| {k} | {d[k][0]} | {d[k][1]} |\n" for k in KEYS if k in d)
    st = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(d[k][0]))
          for k in d if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k]
    stalls = ", ".join("%s %.2f" % kv for kv in sorted(st, key=lambda x: -x[1])[:8])
    traffic = bytes_of(d, "dram__bytes_read.sum") + bytes_of(d, "dram__bytes_write.sum")
    lines = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, "12", "--stalls"], capture_output=True, text=True).stdout
    md = f"## {title}\n\n{what}\n\n{tab}\nWarp stall reasons (warps stalled per issue-active cycle): {stalls}\n\n"
    md += f"DRAM traffic of the launch: {traffic / 1e9:.3f} GB"
    if alg_bytes:
        md += f" against {alg_bytes / 1e9:.3f} GB of algorithmic bytes (ratio {traffic / alg_bytes:.3f})"
    md += ".\n\nSource lines with the most stall samples:\n\n```\n" + "\n".join(lines.splitlines()[-13:]) + "\n```\n\n"
    return md, traffic


for f in ("r02_bench_line", "r02_bench_reference_line", "r02_bench_line_c2", "r02_bench_line_c5", "r02_bench_2gpu", "r02_bench_2gpu_c5"):
    src = f"gpurun_out/{f}.json"
    if os.path.exists(src):
        shutil.copy(src, f"profiles/{f}.json" if not f.startswith("r02_bench_2gpu") else "profiles/" + f.replace("r02_bench_2gpu", "r02_bench_line_2gpu") + ".json")
shutil.copy("gpurun_out/r02_launches_bench.csv", "profiles/r02_launches_bench.csv")
def launch_list(path):
    """mean microseconds per kernel name of the last full step in the launch list"""
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    per = {}
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        name = name.split("<")[0]
        per.setdefault(name, []).append(float(r[iv].replace(",", "")) / 1e3)
    mean = {k: sum(v) / len(v) for k, v in per.items()}
    steps = len(per.get("airs_plan_kernel", [1]))
    out = {k: v for k, v in mean.items()}
    out["total"] = sum(sum(v) for v in per.values()) / steps
    idle = [v for k, v in mean.items() if k not in ("airs_plan_kernel", "airs_fast_kernel")]
    out["idle_min"], out["idle_max"] = (min(idle), max(idle)) if idle else (0.0, 0.0)
    return out


launch = launch_list("gpurun_out/r02_launches_bench.csv")
bl = json.load(open("profiles/r02_bench_line.json"))
b2 = json.load(open("profiles/r02_bench_line_c2.json"))
head = sys.argv[1] if len(sys.argv) > 1 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
md = f"""# Round 2 - ncu --set full summaries of the three encode kernels and the transform kernel

Captured on a B200 with `ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1`
(tools/gpu_call.sh steps ncubench / ncucase), each after the same command had exited 0 without ncu.  Numbers taken
under ncu are never bench values; the bench lines of the same code are profiles/r02_bench_line*.json (default line:
{bl['value']:.0f} GB/s input device-timed, roofline frac {bl['roofline']['frac']:.3f}, e2e {bl['e2e']['value']:.1f} GB/s,
CPU reference {bl['cpu_baseline']['value']:.2f} GB/s on {bl['cpu_baseline']['cores']} cores, parity_full {bl['parity_full']}).

Launch list of the default bench command (`--metrics gpu__time_duration.sum -k regex:airs|concat`,
profiles/r02_launches_bench.csv; cold and serialised under ncu): per step one airs_plan_kernel ({launch['airs_plan_kernel']:.0f} us for the
1 Mi jobs), one airs_fast_kernel ({launch['airs_fast_kernel'] / 1e3:.2f} ms) and the kernels that find nothing to do in this workload (both
airs_tile_kernel variants, airs_raw_kernel, airs_encode_kernel, airs_checksum_kernel: {launch['idle_min']:.0f}-{launch['idle_max']:.0f} us each; no work buffers: the
transform kernels are not launched) - airs_fast_kernel is {100 * launch['airs_fast_kernel'] / launch['total']:.1f} % of the listed time and
{100 * launch['airs_fast_kernel'] / 1e3 / bl['ms_per_step']:.1f} % of the device-timed step ({bl['ms_per_step']:.3f} ms): the share agrees.

"""
t3 = None
s, t3 = section("airs_fast_kernel on the default workload (config 3: 1 Mi chunks of 4 KiB, escape heavy)", "gpurun_out/r02_ncu_fast_c3.ncu-rep",
                "`python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity --no-extras`, fourth launch of the kernel.",
                bl["roofline"]["algorithmic_bytes_per_launch"])
md += s
s, t2 = section("airs_encode_kernel on config 2 batched (888 contexts x 256 frames x 64 KiB)", "gpurun_out/r02_ncu_encode_c2b.ncu-rep",
                "`python bench.py --workload c2 ...`, fourth launch of the kernel.", b2["roofline"]["algorithmic_bytes_per_launch"])
md += s
s, _ = section("airs_tile_kernel<false> on 256 chunks of 4 MiB (probe c256)", "gpurun_out/r02_ncu_tile_c256.ncu-rep",
               "`python tools/ncu_case.py c256`, second launch of the variant for single frames (1 GiB of samples, 524288 tiles).")
md += s
if os.path.exists("gpurun_out/r02_ncu_iwt.ncu-rep"):
    s, _ = section("airs_iwt_kernel on 16384 IWT frames of 64 KiB (probe iwt32k)", "gpurun_out/r02_ncu_iwt.ncu-rep",
                   "`python tools/ncu_case.py iwt32k`, second launch (1 GiB of samples in, 1 GiB of coefficients out).", 2 * (1 << 30))
    md += s
open("profiles/r02_ncu_summary.md", "w").write(md)
json.dump({"c3": {"units": 1 << 20, "dram_bytes_per_launch": t3, "source": "profiles/r02_ncu_summary.md (ncu --set full of airs_fast_kernel, bench.py default workload), code of commit " + head},
           "c2": {"units": int(b2["config"]["per_gpu_input_bytes"] // (256 * 65536)), "dram_bytes_per_launch": t2, "source": "profiles/r02_ncu_summary.md (ncu --set full of airs_encode_kernel, bench.py --workload c2), code of commit " + head}},
          open("profiles/roofline_traffic.json", "w"))
print("written")
