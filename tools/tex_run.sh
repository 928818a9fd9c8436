mkdir -p gpurun_out/ref4
oracle/_ref/ref_dump render scenes/xml/medium_cbox_vpt_256.xml gpurun_out/ref4/medium_cbox_256 vpt 0 1,2,16,64 > gpurun_out/ref4/log.txt 2>&1
tail -5 gpurun_out/ref4/log.txt; ls -la gpurun_out/ref4
oracle/_ref/ref_dump bench scenes/xml/medium_cbox_vpt_256.xml vpt 4 32 2>/dev/null | grep ref_bench
