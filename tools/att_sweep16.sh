B=tools/bin/att_bench; L=vfmseg_b200/lib/libvfmseg_b200.so
for a in "1 128 1" "2 256 2" "1 1024 2" "3 1025 4" "2 197 2" "1 2049 2" "2 17 2" "1 385 3" "2 641 2" "20 257 12" "40 65 8" "600 17 1" "300 129 2" "7 1025 16"; do echo "== $a"; timeout 60 $B $L $a 6,7 2; done
echo "== peaked"; timeout 60 $B $L 2 1025 3 6,7 2 3.0
echo "== 2 2 1"; timeout 60 $B $L 2 2 1 7 2
echo "---- timing"; timeout 100 $B $L 36 1025 16 1,5,6,7 20; timeout 100 $B $L 18 1025 16 1,5,7 20
