"""Synthetic workloads (test / bench infrastructure -- NOT part of the engine).

cohort.py    seeded proteome, variant catalogue and phased cohorts -> Task batches on the host (numpy), the host twin
             of the device task generator and the bit-exact specification the GPU tests hold it to
devgen.py    the same cohorts' per-haplotype site lists produced ON the GPU (synth/devgen.cu -> libv2p_synth.so), so
             that a 50,000-sample cohort can be streamed without 300 s of numpy; a numpy twin pins it
Nothing under vcf2prot_b200/ imports this package."""
