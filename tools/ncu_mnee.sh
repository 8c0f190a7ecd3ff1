out=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_mnee" -s 4 -c 2 -f -o /tmp/mnee python tools/kernel_times.py grating 1920 1080 1 > $out/r1_mnee_ncu.log 2>&1
ncu -i /tmp/mnee.ncu-rep --page raw --csv > $out/r1_mnee_raw.csv 2>>$out/r1_mnee_ncu.log
ncu -i /tmp/mnee.ncu-rep --page source --csv 2>>$out/r1_mnee_ncu.log | gzip -9 > $out/r1_mnee_source.csv.gz
ncu -i /tmp/mnee.ncu-rep --page source --print-source cuda,sass --csv 2>>$out/r1_mnee_ncu.log | gzip -9 > $out/r1_mnee_cuda.csv.gz
tail -3 $out/r1_mnee_ncu.log
