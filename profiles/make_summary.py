#!/usr/bin/env python
"""Tables for profiles/*_SUMMARY.md from files a gpurun call brought back.

    python profiles/make_summary.py launches <ncu launch list .csv> <bench .json of the same command>
        -> stage shares: ncu per-launch times (cold cache, serialised) against the CUDA-event stage timers
    python profiles/make_summary.py raw <ncu --page raw --csv> [...]
        -> one row of headline metrics per captured kernel
"""
import csv
import json
import re
import sys
from collections import defaultdict

STAGE_OF = [  # kernel name prefix -> stage label used by vp_stage (csrc/*.cu); round-2 kernels first, round-1 names kept
    ("k_bin_hist", "k1a_bin_hist"), ("k_bin_offsets", "k1a_bin_hist"), ("k_bin_scatter", "k1b_bin_scatter"),
    ("k_cell_count", "k1c_cell_count"), ("k_scan_", "k1d_cell_scan"), ("k_cell_place", "k1e_cell_place"),
    ("k_search_brick", "k1f_search_brick"), ("k_search_crowded", "k1f_search_brick"), ("k_search_rows", "k1f_search_rows"),
    ("k_search_block4", "k1g_search_block4"), ("k_search_exact", "k1h_search_exact"),
    ("k_fft_x_pow", "k4c_fft_x_pow"), ("k_bin_tiles", "k5_bin_tiles"),
    ("k_keygen_pack", "k1a_keygen_pack"), ("k_tile_hist", "k1b_radix_sort"),
    ("k_scatter", "k1b_radix_sort"), ("k_row_starts", "k1d_row_starts"), ("k_group_rows", "k1c_group_rows"),
    ("k_permute", "k1c_permute"), ("k_search_block2", "k1e_search_block2"),
    ("k_fields_sorted", "k3_fields_sorted"), ("k_fft_z", "k4a_fft_z"),
    ("k_fft_y", "k4b_fft_y"), ("k_fft_x_bin", "k4c_fft_x_bin"), ("k_plane_bin", "k5_plane_bin"),
]


def short(name):
    m = re.search(r"(k_[a-z0-9_]+)", name)
    return m.group(1) if m else name


def launches(csv_path, bench_path):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10 and r[0].isdigit()]
    per_kernel, count = defaultdict(float), defaultdict(int)
    for r in rows:
        k = short(r[4])
        per_kernel[k] += float(r[-1]) / 1e3          # ns -> us
        count[k] += 1
    per_stage = defaultdict(float)
    for k, us in per_kernel.items():
        for pre, st in STAGE_OF:
            if k.startswith(pre):
                per_stage[st] += us
                break
    bench = json.load(open(bench_path))
    ev = {k: v["ms_per_step"] for k, v in bench["stages"].items()}
    tot_n, tot_e = sum(per_stage.values()), sum(ev.values())
    print("| stage | ncu total us | ncu share | CUDA-event ms/step | event share |\n|---|---|---|---|---|")
    for st, us in sorted(per_stage.items(), key=lambda kv: -kv[1]):
        e = ev.get(st, float("nan"))
        print(f"| {st} | {us:.1f} | {100 * us / tot_n:.1f}% | {e:.4f} | {100 * e / tot_e:.1f}% |")
    print("\n| kernel | launches | total us |\n|---|---|---|")
    for k, us in sorted(per_kernel.items(), key=lambda kv: -kv[1]):
        print(f"| {k} | {count[k]} | {us:.1f} |")


def raw(paths):
    want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
            ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %peak"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"),
            ("smsp__issue_active.avg.per_cycle_active", "issue/cycle/SMSP"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("launch__registers_per_thread", "regs"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %")]
    print("| kernel | " + " | ".join(w[1] for w in want) + " |\n|" + "---|" * (len(want) + 1))
    for p in paths:
        rows = list(csv.reader(open(p)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            cells = []
            for key, _ in want:
                cells.append(f"{r[ix[key]]} {units[ix[key]]}".strip() if key in ix else "-")
            print(f"| {short(r[ix['Kernel Name']])} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        raw(sys.argv[2:])
