"""Drop-in mirrors of the reference's entry points for the hot path:

  classification  UnlearnMethod / SFRon / SalUn / create_unlearn_method   (Classification/unlearn/)
  masks           generate_fisher_mask (DDPM, SD) and generate_mask (DiT) CLIs, same flags and files
  diffusion       Fisher / top-k mask / per-sample FIM / SFR-on and SalUn forget loops of the DDPM runner,
                  the DiT scripts and the SD train-scripts, over caller-supplied loss closures
"""
from .classification import SFRon, SalUn, UnlearnMethod, create_unlearn_method  # noqa: F401
