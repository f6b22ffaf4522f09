#!/bin/bash
# Python-free GPU smoke of the file-level features through the CLI binaries (seconds, no torch import):
#   gpurun --timeout 120 -- 'bash tests/tools_quick_cli_check.sh'
# Every step runs under `timeout`; results go to gpurun_out/quick_cli.txt.
G=tests/golden; O=gpurun_out/q; mkdir -p $O; R=gpurun_out/quick_cli.txt; : > $R
E=lepton_b200/bin/lepton-b200
say() { echo "$@" | tee -a $R; }
same() { cmp -s "$1" "$2" && echo same || echo DIFFERENT; }
timeout 40 $E -skipverify -outdir=$O $G/androidcrop.jpg $G/grayscale.jpg $G/iphoneprogressive.jpg $G/trailingrst.jpg $G/android.lep $G/gray2sf.lep $G/iphonecrop2_t8.lep 2>>$R; say "batch rc=$?"
for n in androidcrop grayscale iphoneprogressive trailingrst; do say " $n.lep $(same $O/$n.lep $G/$n.lep)"; done
say " android.jpg $(same $O/android.jpg $G/android.jpg)  gray2sf.jpg $(same $O/gray2sf.jpg $G/gray2sf.jpg)  iphonecrop2_t8.jpg $(same $O/iphonecrop2_t8.jpg $G/iphonecrop2.jpg)"
timeout 30 $E $G/legacy/gold-legacy.lep $O/legacy.jpg 2>>$R; say "legacy rc=$? md5=$(md5sum < $O/legacy.jpg | cut -c1-32) (want 9ffbfc24d1157d0b1ed7a9b53bef4c23)"
timeout 30 $E $G/legacy/roundtripfail.jpg $O/rf.lep 2>>$R; say "verify on roundtripfail rc=$? (want 41) output exists: $(test -e $O/rf.lep && echo yes || echo no)"
timeout 30 $E $G/androidcrop.jpg $O/v.lep 2>>$R; say "verify on androidcrop rc=$? $(same $O/v.lep $G/androidcrop.lep)"
timeout 30 $E -skipverify -minencodethreads=4 $G/android.jpg $O/t4.lep 2>>$R; say "minencodethreads=4 rc=$? $(same $O/t4.lep $G/android_t4.lep)"
timeout 30 $E -skipverify -rejectprogressive $G/iphoneprogressive.jpg $O/rp.lep 2>>$R; say "rejectprogressive rc=$? (want 8)"
mkdir -p $O/m; timeout 40 $E -skipverify -devices=0,0 -outdir=$O/m $G/androidcrop.jpg $G/grayscale.jpg $G/iphonecrop2.jpg $G/android.lep 2>>$R; say "two codecs rc=$? $(same $O/m/androidcrop.lep $G/androidcrop.lep) $(same $O/m/iphonecrop2.lep $G/iphonecrop2.lep) $(same $O/m/android.jpg $G/android.jpg)"
P=oracle/_ref/lepton-b200plug
if [ -x $P ]; then
  timeout 40 $P -unjailed -skipverify $G/androidcrop.jpg $O/p.lep >/dev/null 2>>$R; say "plug encode rc=$? $(same $O/p.lep $G/androidcrop.lep)"
  timeout 40 $P -unjailed -forceprogressive $G/androidcrop.lep $O/p1.jpg >/dev/null 2>>$R; say "plug decode (full planes) rc=$? $(same $O/p1.jpg $G/androidcrop.jpg)"
  timeout 40 $P -unjailed $G/androidcrop.lep $O/p2.jpg >/dev/null 2>>$R; say "plug decode (rows) rc=$? $(same $O/p2.jpg $G/androidcrop.jpg)"
fi
rm -rf $O
