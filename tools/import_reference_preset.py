#!/usr/bin/env python3
"""Extract the reference's shipped parameters into poreplex_b200/presets/.

Run once in the build container (the reference tree is not present on the GPU box):

    python tools/import_reference_preset.py [/root/reference]

Writes
  presets/rna_r941.json                    the YAML preset (presets/rna-r941.cfg) as JSON
  presets/MIN106-RNA001/scaler-r3.npz      weights + input_defs/output_transform attrs
  presets/MIN106-RNA001/demux-tetra-r4.npz weights + calibration table
and records the sha256 of each source file so that a later reader can tell which
parameter set it holds.  Only parameter DATA is extracted; no reference source code.
"""
import hashlib
import json
import os
import sys

import yaml

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from poreplex_b200 import params  # noqa: E402


def sha256(path):
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()


def main(ref='/root/reference'):
    src = os.path.join(ref, 'poreplex', 'presets')
    dst = params.PRESET_DIR
    os.makedirs(os.path.join(dst, 'MIN106-RNA001'), exist_ok=True)

    cfg_path = os.path.join(src, 'rna-r941.cfg')
    with open(cfg_path) as f:
        preset = yaml.safe_load(f)
    provenance = {'rna-r941.cfg': sha256(cfg_path)}

    sc_path = os.path.join(src, preset['signal_processing']['scaler_model'])
    dm_path = os.path.join(src, preset['demultiplexing']['demux_model'])
    provenance[os.path.basename(sc_path)] = sha256(sc_path)
    provenance[os.path.basename(dm_path)] = sha256(dm_path)
    params.save_scaler_npz(params.load_scaler_model(sc_path),
                           os.path.join(dst, 'MIN106-RNA001', 'scaler-r3.npz'))
    params.save_demux_npz(params.load_demux_model(dm_path),
                          os.path.join(dst, 'MIN106-RNA001', 'demux-tetra-r4.npz'))

    # kmersize = len(kmermodel.index[0]) (worker_persistence.py:64-66); the k-mer model
    # is a git submodule that is only used for this one constant.
    preset['kmersize'] = 5
    preset['_provenance'] = provenance
    with open(os.path.join(dst, 'rna_r941.json'), 'w') as f:
        json.dump(preset, f, indent=1)
    print('wrote', dst)


if __name__ == '__main__':
    main(*sys.argv[1:])
