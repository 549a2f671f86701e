#!/bin/bash
# build the library from anywhere; TRACE=1 adds the chain-kernel clock stamps
cd /root/repo
if [ -n "$TRACE" ]; then LDM_CHAIN_TRACE=1 python oxford-102-flower-gan-vae-latent-diffusion_b200/build.py --force | tail -1
else python oxford-102-flower-gan-vae-latent-diffusion_b200/build.py "$@" | tail -1; fi
grep -A3 "chain_kernelILi3" oxford-102-flower-gan-vae-latent-diffusion_b200/build/chain.o.ptxas.log | grep "Used\|spill" | tr '\n' ' '; echo
