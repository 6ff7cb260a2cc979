"""Synthetic inputs moved to synthetic/pages.py (they are data generators, not oracle code); re-exported here."""
from synthetic.pages import *  # noqa: F401,F403
from synthetic.pages import INK_BGR, dense_page, random_score_maps, score_maps_from_page, synth_page  # noqa: F401
