"""DRAM traffic per kernel class of one training step, from an ncu launch list taken with
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
       --log-file launches.csv python tools/profile_step.py --batch 64
Classes follow bench.py's `kernel_classes` where a class is identifiable by kernel name (forward and data-gradient convolutions share their kernels
and are reported together).  Writes profiles/roofline_traffic.json: {class: bytes per step}; bench.py copies the dominant class into roofline.traffic."""
import csv, json, re, sys
from collections import defaultdict

CLASSES = [("bn_silu_bwd", r"bn_silu_bwd_(reduce|apply)_kernel|bn_param_grad"),
           ("bn_apply_silu (+ finalize)", r"bn_apply_silu_kernel|bn_finalize_kernel"),
           ("wgrad (wgrad_gemm + reduce)", r"wgrad_(gemm|reduce)_kernel"),
           ("conv_fwd + dgrad + pred convs (conv_gemm kernels)", r"conv_gemm_"),
           ("spp_pool", r"spp_pool_tiled_kernel|spp_pool_kernel"),
           ("spp_pool_bwd", r"spp_pool_bwd"),
           ("simota_assign", r"simota_"),
           ("yolox_loss", r"yolox_loss"),
           ("pack_weights", r"pack_conv_weight"),
           ("preprocess_focus", r"preprocess_focus")]
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if not l.startswith("==")]
per = defaultdict(lambda: defaultdict(float))
units = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}
for r in csv.DictReader(lines):
    name = r["Kernel Name"]
    cls = next((c for c, pat in CLASSES if re.search(pat, name)), "other")
    v = float(r["Metric Value"].replace(",", "")) * units.get(r.get("Metric Unit", ""), 1)
    per[cls][r["Metric Name"]] += v
    if r["Metric Name"] == "gpu__time_duration.sum":
        per[cls]["launches"] += 1
out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of each kernel class in ONE bs=64 640x640 training step "
                   "(ncu, serialised, cold: compare with bench.py's algorithmic bytes of the class, not with its timing); tools/class_traffic.py"}
print("| class | launches | ncu ms | DRAM read MB | DRAM write MB |\n|---|---:|---:|---:|---:|")
for cls, m in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
    out[cls] = int(rd + wr)
    print("| %s | %d | %.3f | %.1f | %.1f |" % (cls, m["launches"], m["gpu__time_duration.sum"] / 1e6, rd / 1e6, wr / 1e6))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
