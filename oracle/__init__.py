"""CPU oracle for the fLDR-VFI custom-kernel hot path (softmax splat + 9x9 correlation).

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``fldr-vfi_b200/``) never imports from here and
fails loudly when its CUDA library is missing.

Contents
--------
``splat_oracle``  torch-CPU restatement of ``softSplat.py`` (reference file:line cited per function)
``corr_oracle``   torch-CPU restatement of ``OpticalFlow/correlation.py``
``build_ref``     recipe that compiles the reference's OWN kernel text (read from
                  ``/root/reference`` where it lies, never copied into git) for the host
                  into ``oracle/_ref/`` (git-ignored)
``ref_host``      ctypes driver for ``oracle/_ref/libref_host.so``
``synth``         seeded synthetic inputs

Parity pin: the reference ships no tests / golden vectors for this path (SURVEY.md
section 4), so the restatement is pinned against the reference's own kernel text executed
on the host (``oracle/_ref``) - see ``tests/golden/make_golden.py`` and
``tests/test_oracle_vs_golden.py``.
"""
