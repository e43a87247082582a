mkdir -p gpurun_out
for a in resnet ecapa; do timeout 600 python scripts/grad_profile.py $a fp32 4 > gpurun_out/grad_profile_${a}_fp32.txt 2>&1; done
head -60 gpurun_out/grad_profile_resnet_fp32.txt
head -8 gpurun_out/grad_profile_ecapa_fp32.txt; sort -k2 -g -r gpurun_out/grad_profile_ecapa_fp32.txt | head -12
