"""scoary_b200 -- B200-native engine for Scoary's per-gene statistics, pairwise
comparisons and permutation path, behind the reference's own Python interface.

Layout:
  csrc/            CUDA kernels (sm_100a) + the C-ABI  -> libscoary_b200.so
  _lib.py          ctypes binding of include/scoary_b200.h (fails loudly without the .so / a GPU)
  engine.py        array-level host API (bit packing, tree flattening, sharding)
  tree.py          nested-list trees: flatten, prune, parse/write
  methods.py       mirror of the reference's call sites (Setup_results, PairWiseComparisons,
                   ConvertUPGMAtoPhyloTree, Permute, StoreResults, main / CLI)
  synth.py         synthetic workloads of SURVEY.md 8(d)
"""
__version__ = "0.1.0"
