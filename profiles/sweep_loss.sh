# A/B of loss-kernel builds / knobs on cfg3: bash profiles/sweep_loss.sh
for v in "" _g256x1 _g256x2 _g128x4 _g128x1; do
  lib=$PWD/keras-object-detection_b200/yolohot/libyolohot$v.so
  [ -f "$lib" ] || continue
  echo "== gather build '$v'"; YH_LIB_PATH=$lib YH_LOSS_GATHER=1 python profiles/prof_loss.py 2>&1 | head -2
done
echo "== ring (default)"; python profiles/prof_loss.py | tail -3
