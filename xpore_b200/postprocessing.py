"""``xpore postprocessing``: keep, per k-mer, only the rows of ``diffmod.table`` whose
``mod_assignment`` is the majority direction of that k-mer.

Behaviour follows ``xpore/scripts/postprocessing.py:3-39`` of the reference:

* the output is ``<diffmod_dir>/majority_direction_kmer_diffmod.table`` with the same header;
* rows are split on ``,`` (no CSV quoting), directions are counted per k-mer over the whole table;
* a k-mer seen with both directions keeps ``higher`` only if it is strictly more frequent than
  ``lower`` (a tie keeps ``lower``, reference line 24); a k-mer seen with one direction keeps it;
* kept rows are copied byte for byte, in input order.

This is a host-side text filter; nothing here touches the GPU.
"""
from __future__ import annotations

import os


def run_postprocessing(diffmod_table_path: str, out_dir: str) -> str:
    with open(diffmod_table_path, "r") as f:
        header = f.readline()
        entries = f.readlines()
    cols = header.strip().split(",")
    kmer_ind, dir_ind = cols.index("kmer"), cols.index("mod_assignment")
    counts = {}
    keys = []
    for ln in entries:
        fields = ln.strip().split(",")
        key = (fields[kmer_ind], fields[dir_ind])
        keys.append(key)
        per = counts.setdefault(key[0], {})
        per[key[1]] = per.get(key[1], 0) + 1
    choose = {}
    for kmer, per in counts.items():
        if len(per) > 1:
            # the reference indexes ['higher'] and ['lower'] directly: any other label next to them raises
            choose[kmer] = "lower" if per["higher"] <= per["lower"] else "higher"
        else:
            choose[kmer] = next(iter(per))
    out_path = os.path.join(out_dir, "majority_direction_kmer_diffmod.table")
    with open(out_path, "w") as out:
        out.write(header)
        for ln, (kmer, direction) in zip(entries, keys):
            if direction == choose[kmer]:
                out.write(ln)
    return out_path


def postprocessing(args) -> str:
    """Entry point of the ``postprocessing`` sub-command (``--diffmod_dir``)."""
    diffmod_dir = args.diffmod_dir
    return run_postprocessing(os.path.join(diffmod_dir, "diffmod.table"), diffmod_dir)
