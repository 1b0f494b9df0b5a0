"""Parameter dataclasses of the SR-artifact simulators, field-for-field the ones the reference's
YAML trees construct (``fetalsyngen/generator/artifacts/utils.py:10-78``)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class ScannerParams:
    resolution_slice_fac_min: float
    resolution_slice_fac_max: float
    resolution_slice_max: int
    slice_thickness_min: float
    slice_thickness_max: float
    gap_min: float
    gap_max: float
    min_num_stack: int
    max_num_stack: int
    max_num_slices: int
    noise_sigma_min: float
    noise_sigma_max: float
    TR_min: float
    TR_max: float
    prob_void: float
    prob_gamma: float
    gamma_std: float
    slice_size: int
    restrict_transform: bool
    txy: float
    resolution_recon: float = None
    slice_noise_threshold: float = 0.1


@dataclass
class StructNoiseMergeParams:
    merge_type: str
    gauss_nloc_min: int = None
    gauss_nloc_max: int = None
    gauss_sigma_mu: float = None
    gauss_sigma_std: float = None
    perlin_res_list: list = None
    perlin_octaves_list: list = None
    perlin_persistence: float = None
    perlin_lacunarity: int = None
    perlin_increase_size: float = None


@dataclass
class ReconMergeParams:
    merge_type: str
    gauss_ngaussians_min: int = None
    gauss_ngaussians_max: int = None
    perlin_res_list: list = None
    perlin_octaves_list: list = None
    perlin_persistence: float = None
    perlin_lacunarity: int = None
    perlin_increase_size: float = None


@dataclass
class ReconParams:
    prob_misreg_slice: float
    slices_misreg_ratio: float
    prob_misreg_stack: float
    txy: float
    prob_smooth: float
    prob_rm_slices: float
    rm_slices_min: float
    rm_slices_max: float
    prob_merge: float
    merge_params: ReconMergeParams
