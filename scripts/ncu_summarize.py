"""Summarise an `ncu --set full` report (.ncu-rep, read here with `ncu -i ... --page raw --csv`) into the per-launch JSON
kept under profiles/: duration, DRAM bytes, tensor-pipe and SM throughput, L2 hit rate, occupancy, instruction count.
Usage: python scripts/ncu_summarize.py gpurun_out/r01_full_v4.ncu-rep profiles/r01_ncu_full_summary.json "<command>"
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

# role of the i-th captured launch of one cfg1 train step, in stream order of the filter
# regex:"gemm_pair_kernel|attn_fwd_pipe_kernel|attn_bwd_kernel"
ROLES = ["GRU x-projection gates [7168x2048x304]", "GRU x-projection candidate [7168x1024x304]",
         "v_linear_v forward [18432x1024x2048], 77.3 GF", "attn_fwd (pipelined)", "attn_bwd",
         "v_linear_v wgrad [2048x1024x18432], 77.3 GF, split-K", "GRU wgrad", "GRU wgrad", "GRU wgrad", "GRU wgrad",
         "dE", "dE"]


def main():
    rep, out, command = sys.argv[1], sys.argv[2], sys.argv[3]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(header)}
    kernels = []
    for i, r in enumerate(data):
        k = {"id": i, "kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        for m in KEEP:
            if m in col and r[col[m]] != "":
                k[m] = float(r[col[m]].replace(",", ""))
                k[m + "__unit"] = units[col[m]]
        rd = k.get("dram__bytes_read.sum", 0.0) * TO_BYTES.get(k.get("dram__bytes_read.sum__unit", "byte"), 1.0)
        wr = k.get("dram__bytes_write.sum", 0.0) * TO_BYTES.get(k.get("dram__bytes_write.sum__unit", "byte"), 1.0)
        k["dram_traffic_bytes"] = rd + wr
        k["time_us"] = k.get("gpu__time_duration.sum", 0.0) * TO_US.get(k.get("gpu__time_duration.sum__unit", "us"), 1.0)
        kernels.append(k)
    # name the launches: pair GEMMs by shape order within the step, attention kernels by name
    gi = 0
    for k in kernels:
        if "attn_fwd" in k["kernel"]:
            k["role"] = "attn_fwd (pipelined)"
        elif "attn_bwd" in k["kernel"]:
            k["role"] = "attn_bwd"
        else:
            k["role"] = f"pair GEMM #{gi}"
            gi += 1
    # the two 77.3 GF launches are the longest pair GEMMs: forward first, wgrad second in stream order
    big = sorted([k for k in kernels if k["role"].startswith("pair GEMM")], key=lambda k: -k["time_us"])[:2]
    big.sort(key=lambda k: k["id"])
    if len(big) == 2:
        big[0]["role"] = "v_linear_v forward [18432x1024x2048], 77.3 GF"
        big[1]["role"] = "v_linear_v wgrad [2048x1024x18432], 77.3 GF, split-K"
    json.dump({"command": command, "note": "one train step (cfg1, bf16 mode); per-launch values; times are cold-cache "
               "and serialised", "kernels": kernels}, open(out, "w"), indent=1)
    for k in kernels:
        print(f'{k["id"]:3d} {k["time_us"]:9.1f} us  {k["dram_traffic_bytes"] / 1e6:8.1f} MB  {k["role"]:55s} {k["kernel"][:60]}')


if __name__ == "__main__":
    main()
