"""ncu CSV (launch list of tools/gpu_probe_craft.py, one craft_forward) -> markdown per-layer table.
usage: python tools/ncu_craft_table.py gpurun_out/ncu_craft.csv NPAGES > profiles/r02_craft_layers_ncu.md"""
import csv, sys
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
rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
hdr = next(r for r in csv.reader(open(path)) if r and r[0] == "ID")
def col(name): return hdr.index(name)
rows = rows[-len(LAYERS):]                       # the last craft_forward of the run
H, W = 2560, 1984
print(f"| layer | kernel | ms ({npages} pages) | real GFLOP/page | TFLOP/s (real) | tensor pipe active % | dram GB |")
print("|---|---|---|---|---|---|---|")
tot_ms = tot_f = 0.0
for (name, cin, cout, taps, div), r in zip(LAYERS, rows):
    ms = float(r[col("gpu__time_duration.sum")]) / 1e6 if float(r[col("gpu__time_duration.sum")]) > 1e4 else float(r[col("gpu__time_duration.sum")])
    gf = 2.0 * cin * cout * taps * (H // div) * (W // div) / 1e9
    kern = r[col("Kernel Name")].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "")[:40]
    tp = r[col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")] if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in hdr else ""
    dr = float(r[col("dram__bytes_read.sum")]) + float(r[col("dram__bytes_write.sum")]) if "dram__bytes_read.sum" in hdr else 0
    tot_ms += ms; tot_f += gf * npages
    print(f"| {name} | `{kern}` | {ms:.3f} | {gf:.1f} | {gf * npages / ms / 1e3 if gf else 0:.0f} | {tp} | {dr / 1e9 if dr > 1e3 else dr:.2f} |")
print(f"| **total** | | {tot_ms:.2f} | {tot_f / npages:.0f} | {tot_f / tot_ms / 1e3:.0f} | | |")
