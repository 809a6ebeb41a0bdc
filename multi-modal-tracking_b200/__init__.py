"""B200-native per-frame forward of the MixViT RGB / RGB-T trackers (drop-in behind the reference's
`lib/models` builders and `lib/test/tracker` classes).  Import as `mmt_b200` (see /mmt_b200.py: the
directory name contains hyphens, so the root-level shim registers it under an importable name).
"""
__version__ = "0.1.0"
