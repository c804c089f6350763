"""Minimal stand-in for the one kornia symbol the reference hot path imports
(NetWorks/PixelShuffleUpsample.py:5). kornia is not installed in this image; test infrastructure only."""
