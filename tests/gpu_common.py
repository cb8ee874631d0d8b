"""Helpers shared by the GPU parity tests: build the product model for a golden case."""
import numpy as np

from helpers import roi_of


def product_model(case, path="auto"):
    import umpa_b200
    cls = {"NoDF": umpa_b200.UMPAModelNoDF, "DF": umpa_b200.UMPAModelDF,
           "DFKernel": umpa_b200.UMPAModelDFKernel}[case["kind"]]
    m = cls(case["sam"], case["ref"], mask_list=case["mask"],
            pos_list=None if case["pos"] is None else [np.array(p) for p in case["pos"]],
            window_size=case["Nw"], max_shift=case["max_shift"])
    if case["assign"] is not None:
        m.assign_coordinates = case["assign"]
    if case["subpx"] is not None:
        m.sub_pixel_mode = case["subpx"]
    m.cuda_path = path
    return m


def run_case(case, path="auto"):
    m = product_model(case, path)
    kw = {}
    if case["step"] is not None:
        kw["step"] = case["step"]
    if case["ROI"] is not None:
        kw["ROI"] = case["ROI"]
    if case["dxdy"] is not None:
        kw["dxdy"] = case["dxdy"]
    if case["kind"] == "DFKernel":
        kw["abc"] = case["abc"]
    res = m.match(quiet=True, **kw)
    return m, res


TABLE_CASES = ("nodf_clean", "df_clean", "nodf_noisy", "df_noisy", "df_lowcontrast", "df_nw3_ms6",
               "nodf_nw1", "df_subpx0", "df_subpx1", "df_step3", "df_roi", "df_dxdy",
               "df_assign_ref", "dfk_clean", "dfk_nw3_ms5", "dfk_roi_step", "dfk_wide_blur",
               "dfk_assign_ref", "dfk_assign_ref_step")
