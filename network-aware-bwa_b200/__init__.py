"""B200-native `bwa aln` hot path.  Import via importlib:
    bwa = importlib.import_module("network-aware-bwa_b200")
"""
