"""B200-native `bwa aln` hot path behind the reference's own call seam.

The directory name carries a hyphen, so import it with importlib:

    import importlib
    bwa = importlib.import_module("network-aware-bwa_b200")
    bwa.api.init(); bwa.api.load_index(idx); bwa.api.aln_flat(bases, offs, opt)

Modules: api (ctypes binding of include/bwa_gpu.h -> libbwagpu.so, no fallback),
abi (struct mirrors), index (FM-index build/load in the reference's formats),
simulate (seeded genomes and reads), workload (cached genome + index files of the BASELINE.json shapes),
build (nvcc recipe for csrc/).
"""
from . import abi, api, build, index, shard, simulate, workload  # noqa: F401
