"""ncu CSV (long format: one row per launch and metric) of tools/gpu_probe_craft.py -> markdown per-layer table of the
LAST craft_forward in the log.  usage: python tools/ncu_craft_table.py gpurun_out/ncu_craft.csv NPAGES"""
import collections, csv, sys
LAYERS = [("conv1_1", 3, 64, 9, 1), ("conv1_2", 64, 64, 9, 1), ("pool1", 0, 0, 0, 2), ("conv2_1", 64, 128, 9, 2), ("conv2_2", 128, 128, 9, 2),
          ("pool2", 0, 0, 0, 4), ("conv3_1", 128, 256, 9, 4), ("conv3_2", 256, 256, 9, 4), ("conv3_3", 256, 256, 9, 4), ("pool3", 0, 0, 0, 8),
          ("conv4_1", 256, 512, 9, 8), ("conv4_2", 512, 512, 9, 8), ("conv4_3", 512, 512, 9, 8), ("pool4", 0, 0, 0, 16),
          ("conv5_1", 512, 512, 9, 16), ("conv5_2", 512, 512, 9, 16), ("pool5 3x3 s1", 0, 0, 0, 16), ("fc6 (dil 6)", 512, 1024, 9, 16),
          ("fc7", 1024, 1024, 1, 16), ("upconv1a", 1536, 512, 1, 16), ("upconv1b", 512, 256, 9, 16), ("up x2", 0, 0, 0, 8),
          ("upconv2a", 768, 256, 1, 8), ("upconv2b", 256, 128, 9, 8), ("up x2", 0, 0, 0, 4), ("upconv3a", 384, 128, 1, 4),
          ("upconv3b", 128, 64, 9, 4), ("up x2", 0, 0, 0, 2), ("upconv4a", 192, 64, 1, 2), ("upconv4b (32 real)", 64, 32, 9, 2),
          ("cls1 (32->32)", 32, 32, 9, 2), ("cls2 (32->32)", 32, 32, 9, 2), ("cls3 (32->16)", 32, 16, 9, 2), ("cls4 (16->16)", 16, 16, 1, 2),
          ("cls5 (16->2)", 16, 2, 1, 2)]
path, npages = sys.argv[1], int(sys.argv[2])
launches = collections.OrderedDict()
for r in csv.reader(open(path)):
    if len(r) > 14 and r[0].isdigit():
        launches.setdefault(int(r[0]), {"k": r[4]})[r[12]] = float(r[14].replace(",", ""))
items = list(launches.values())
start = [i for i, d in enumerate(items) if "conv1_1" in d["k"]][-1]
H, W = 2560, 1984
print(f"| layer | kernel | ms ({npages} pages) | real GFLOP/page | TFLOP/s (real) | tensor pipe active % | L2 throughput % | DRAM GB | DRAM TB/s |")
print("|---|---|---|---|---|---|---|---|---|")
tot_ms = tot_f = 0.0
for (name, cin, cout, taps, div), d in zip(LAYERS, items[start:]):
    ms = d["gpu__time_duration.sum"] / 1e6
    gf = 2.0 * cin * cout * taps * (H // div) * (W // div) / 1e9
    kern = d["k"].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "")
    dram = (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / 1e9
    tot_ms += ms
    tot_f += gf * npages
    print(f"| {name} | `{kern}` | {ms:.3f} | {gf:.1f} | {gf * npages / ms:.0f} | {d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} | "
          f"{d.get('lts__throughput.avg.pct_of_peak_sustained_elapsed', 0):.0f} | {dram:.2f} | {dram / ms:.2f} |")
print(f"| **total** | | {tot_ms:.2f} | {tot_f / npages:.0f} | {tot_f / tot_ms:.0f} | | | | |")
