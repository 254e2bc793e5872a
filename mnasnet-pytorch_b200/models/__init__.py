"""Drop-in replacement for the reference's ``src/models`` package (mnasnet + classifiers only): put
``mnasnet-pytorch_b200/`` on ``sys.path`` ahead of the reference's ``src/`` and
``from models.classifiers import load_model, FineTuneModelPool`` (src/train.py:38) resolves here."""
