#!/bin/bash
# Round-start probe (SURVEY §8c "Action every round"): is there any libjxl / .jxl / JXL-capable tool on the GPU box?
out=gpurun_out/probe_libjxl.log
{
echo "== date"; date -u
echo "== nproc"; nproc; lscpu | head -20
echo "== find *jxl* (outside the repo copy)"; find / -xdev \( -path /proc -o -path /sys -o -path "$GRAFT_REPO_ROOT" -o -path /root/repo \) -prune -o -iname '*jxl*' -print 2>/dev/null | head -50
echo "== find *.jxl"; find / -xdev \( -path /proc -o -path /sys -o -path "$GRAFT_REPO_ROOT" -o -path /root/repo \) -prune -o -name '*.jxl' -print 2>/dev/null | head -20
echo "== ldconfig jxl/hwy/brotli/lcms"; ldconfig -p | grep -iE 'jxl|hwy|brotli|lcms' 
echo "== tools"; for t in cjxl djxl benchmark_xl jxlinfo convert magick vips ffmpeg gm; do printf "%s: " $t; command -v $t || echo absent; done
echo "== python bindings"; for m in imagecodecs pillow_jxl jxlpy pyvips imageio pillow_jxl_plugin PIL cv2; do python - <<PY
try:
    import $m
    print("$m: present", getattr($m, "__version__", ""))
except Exception as e:
    print("$m: absent (%s)" % type(e).__name__)
PY
done
python - <<PY
try:
    from PIL import features, Image
    print("PIL codecs:", [f for f in features.get_supported()])
    print("PIL has .jxl:", ".jxl" in Image.registered_extensions())
except Exception as e:
    print("PIL probe failed", e)
try:
    import cv2
    print("cv2 jxl reader:", cv2.haveImageReader("x.jxl"))
except Exception as e:
    print("cv2 probe failed", e)
PY
echo "== baseline/_ref"; ls -la baseline/_ref 2>&1 | head
echo "== /root/reference on box?"; ls /root/reference 2>&1 | head -3
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
} > $out 2>&1
echo probe done
