cp scannertools_b200/libscannertools_b200.so /tmp/current.so
for rep in 1 2; do
for v in h4 h6 h8; do
  cp tools/ab/$v.so scannertools_b200/libscannertools_b200.so
  echo "== $v"
  python tools/hist_batch_probe.py | grep -E "4K n=(16|32|40)|1080p n=64"
done
done
cp /tmp/current.so scannertools_b200/libscannertools_b200.so
