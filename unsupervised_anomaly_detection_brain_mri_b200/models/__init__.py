"""Model functions with the reference's names and call signatures (reference models/*.py).

In the reference these build TF graph nodes; here they build a ``GraphSpec`` - a description of the same graph that
the trainers hand to the CUDA engine.  File name == function name is preserved because run.py looks networks up by it
(reference run.py:22-24)."""
