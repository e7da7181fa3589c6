"""CPU oracle for the ABC-Net hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``abcnet_b200``) imports this directory. Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may use it, and only as the checker / the timed CPU
baseline, never as the shipped path.

Contents (each module cites the reference file:line it follows):

* ``detrand``     -- re-export of ``synthdata.detrand`` (platform-exact deterministic pseudo-random
                     tensors, splitmix64). Plain synthetic inputs / weights (binary images, dense
                     targets, ``make_state_dict``) live in the top-level ``synthdata`` package so that
                     ``bench.py``'s GPU arm and ``tools/`` never import the oracle.
* ``unet_ref``    -- functional fp32 restatement of ``src/unet.py`` on a state_dict.
* ``decode_ref``  -- numpy restatement of ``src/img2smiles.py:62-80,105-193``.
* ``assemble_ref``-- restatement of the host assembly ``src/img2smiles.py:195-318`` and
                     of the MOL-block text of ``src/generate_smiles.py:18-105``.
* ``loss_ref``    -- fp64 restatement of ``src/train.py:95-137``.
* ``synth``       -- synthetic test cases that restate reference rules (planted heat-maps, pseudo-molecule
                     drawings with labels, target rasterisation); re-exports ``synthdata.inputs``.
* ``targets_ref`` -- restatement of the dense-target rasterisation ``src/utils.py:83-228``.

Pinning: the reference ships no golden vectors. The oracle is pinned
against outputs of the reference itself, produced in the build container by
``tests/golden/make_golden.py`` (imports ``/root/reference/src/unet.py`` and executes
source slices of ``img2smiles.py`` / ``train.py`` / ``generate_smiles.py`` in place),
committed under ``tests/golden/``.
"""
