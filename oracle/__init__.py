"""TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference hot path (``FastEditor.edit`` in the reference's
``src/pipeline.py``).  Nothing under ``oracle/`` is imported by the product package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs use it,
and only as the checker / the reported CPU baseline.
"""
