#!/usr/bin/env python
"""profiles/traffic.json from an ncu launch list that carries dram__bytes_read/write.sum per launch
(scripts/gpu_round_g.sh: one bench step under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`).

usage: python scripts/traffic_from_ncu.py profiles/r01_traffic_cfg2_g.csv > profiles/traffic.json
The LAST step in the capture is used (everything from the last k_im2col_rgb launch on). Families follow bench.py's names;
the 13 forward convolutions are the launches right after k_im2col_rgb. Values are DRAM bytes (read + write) per launch."""
import collections
import csv
import json
import sys

VGG = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3",
       "conv5_1", "conv5_2", "conv5_3"]
BY_NAME = {"k_ce(": "ce", "k_adam(": "adam", "k_sumsq(": "sumsq", "k_sample_z(": "sample_z", "k_dz_reduce(": "dz_reduce",
           "lstm_seq_kernel<SeqFwdEpi>": "lstm_fwd_seq", "lstm_seq_kernel<SeqBwdEpi>": "lstm_bwd_seq",
           "lstm_seq2_kernel<Seq2FwdEpi": "lstm_fwd_seq", "lstm_seq2_kernel<Seq2BwdEpi": "lstm_bwd_seq",
           "k_sample_z_v4": "sample_z", "k_dz_reduce_v4": "dz_reduce",
           "k_im2col_rgb": "im2col_rgb", "k_embed_gather(": "embed_gather", "k_embed_scatter(": "embed_scatter",
           "k_colsum_bf16(": "colsum_bf16"}


def main(path):
    launches = collections.OrderedDict()
    for r in csv.reader(open(path)):
        if len(r) > 14 and r[0].isdigit():
            launches.setdefault(int(r[0]), {"name": r[4]})[r[12]] = float(r[14].replace(",", ""))
    ids = sorted(launches)
    start = max(i for i in ids if "k_im2col_rgb" in launches[i]["name"])
    step = [launches[i] for i in ids if i >= start]
    # the capture may run past the end of the step into the next shadow refresh; that does not matter for the families below
    fam = collections.defaultdict(list)
    for j, name in enumerate(VGG):
        L = step[1 + j]
        fam[name].append(L)
        fam["conv"].append(L)
    for L in step:
        for key, f in BY_NAME.items():
            if key in L["name"]:
                fam[f].append(L)
    out = {}
    for f, ls in fam.items():
        out[f] = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in ls) / len(ls)
    out["_source"] = path
    out["_note"] = "DRAM bytes (read + write) per launch, averaged over the family's launches in one cfg2 step"
    json.dump(out, sys.stdout, indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1])
