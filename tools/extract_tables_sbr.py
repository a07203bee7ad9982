"""SBR part of the build-time table extractor (loaded by extract_tables.py).

Same rules as the parent: tables are data read verbatim from the reference sources
(sbr/*.java) and emitted as float32 bit patterns / integers.
"""


def emit(em, ref, JavaFile, AAC):
    S = AAC + "sbr/"
    em.raw("/* ======================= SBR tables (sbr package) ======================= */")
    fb = JavaFile(ref, S + "Filterbank.java")
    em.f32("SBR_QMF_C", fb.floats("qmf_c"), "sbr/Filterbank.java:7  QMF prototype window")
    dct = JavaFile(ref, S + "DCT.java")
    em.f32("SBR_W_ARRAY_REAL", dct.floats("w_array_real"), "sbr/DCT.java:8")
    em.f32("SBR_W_ARRAY_IMAG", dct.floats("w_array_imag"), "sbr/DCT.java:20")
    em.f32("SBR_DCT4_64_TAB", dct.floats("dct4_64_tab"), "sbr/DCT.java:31")
    em.i32("SBR_BIT_REV_TAB", dct.ints("bit_rev_tab"), "int32_t", "sbr/DCT.java:130")
    ne = JavaFile(ref, S + "NoiseEnvelope.java")
    for nm in ("E_deq_tab", "Q_div2_tab", "Q_div2_tab_left", "Q_div2_tab_right", "Q_div_tab", "Q_div_tab_left",
               "Q_div_tab_right", "E_pan_tab"):
        em.f32("SBR_" + nm.upper(), ne.floats(nm), "sbr/NoiseEnvelope.java")
    nt = JavaFile(ref, S + "NoiseTable.java")
    em.f32("SBR_NOISE_TABLE", nt.floats("NOISE_TABLE"), "sbr/NoiseTable.java:6")
    hf = JavaFile(ref, S + "HFAdjustment.java")
    em.f32("SBR_H_SMOOTH", hf.floats("h_smooth"), "sbr/HFAdjustment.java:5")
    em.f32("SBR_LIM_GAIN", hf.floats("limGain"), "sbr/HFAdjustment.java:13")
    hg = JavaFile(ref, S + "HFGeneration.java")
    em.i32("SBR_GOAL_SB_TAB", hg.ints("goalSbTab"), "int32_t", "sbr/HFGeneration.java:5")
    fbt = JavaFile(ref, S + "FBT.java")
    em.i32("SBR_START_MIN_TABLE", fbt.ints("startMinTable"), "int32_t", "sbr/FBT.java:9")
    em.i32("SBR_OFFSET_INDEX_TABLE", fbt.ints("offsetIndexTable"), "int32_t", "sbr/FBT.java:11")
    em.i32("SBR_OFFSET", fbt.ints("OFFSET"), "int32_t", "sbr/FBT.java:13  [7][16]")
    em.i32("SBR_STOP_MIN_TABLE", fbt.ints("stopMinTable"), "int32_t", "sbr/FBT.java:43")
    em.i32("SBR_STOP_OFFSET_TABLE", fbt.ints("STOP_OFFSET_TABLE"), "int32_t", "sbr/FBT.java:46  [12][14]")
    em.f32("SBR_LIMITER_BANDS_COMPARE", fbt.floats("limiterBandsCompare"), "sbr/FBT.java:327")
    ht = JavaFile(ref, S + "HuffmanTables.java")
    for nm in ("T_HUFFMAN_ENV_1_5DB", "F_HUFFMAN_ENV_1_5DB", "T_HUFFMAN_ENV_BAL_1_5DB", "F_HUFFMAN_ENV_BAL_1_5DB",
               "T_HUFFMAN_ENV_3_0DB", "F_HUFFMAN_ENV_3_0DB", "T_HUFFMAN_ENV_BAL_3_0DB", "F_HUFFMAN_ENV_BAL_3_0DB",
               "T_HUFFMAN_NOISE_3_0DB", "T_HUFFMAN_NOISE_BAL_3_0DB"):
        em.i32("SBR_" + nm, ht.ints(nm), "int16_t", "sbr/HuffmanTables.java  binary tree, rows {next0, next1}, leaf = value-64 (<0)")
    emit_ps(em, ref, JavaFile, AAC)


def emit_ps(em, ref, JavaFile, AAC):
    P = AAC + "ps/"
    em.raw("/* ======================= PS tables (ps package) ======================= */")
    t = JavaFile(ref, P + "PSTables.java")
    em.f32("PS_FILTER_A", t.floats("filter_a"), "ps/PSTables.java:20")
    for nm in ("Phi_Fract_Qmf", "Phi_Fract_SubQmf20", "Q_Fract_allpass_Qmf", "Q_Fract_allpass_SubQmf20", "cos_alphas", "sin_alphas",
               "cos_betas_normal", "sin_betas_normal", "cos_betas_fine", "sin_betas_fine", "sincos_alphas_B_normal",
               "sincos_alphas_B_fine", "cos_gammas_normal", "cos_gammas_fine", "sin_gammas_normal", "sin_gammas_fine",
               "sf_iid_normal", "sf_iid_fine", "ipdopd_cos_tab", "ipdopd_sin_tab"):
        em.f32("PS_" + nm.upper(), t.floats(nm), "ps/PSTables.java")
    h = JavaFile(ref, P + "Huffman.java")
    for nm in ("f_huff_iid_def", "t_huff_iid_def", "f_huff_iid_fine", "t_huff_iid_fine", "f_huff_icc", "t_huff_icc",
               "f_huff_ipd", "t_huff_ipd", "f_huff_opd", "t_huff_opd"):
        em.i32("PS_" + nm.upper(), h.ints(nm), "int16_t", "ps/Huffman.java  binary tree, rows {next0, next1}, leaf = value-31 (<0)")
    em.f32("PS_P2_13_20", JavaFile(ref, P + "Filter2.java").floats("p2_13_20"), "ps/Filter2.java:15")
    em.f32("PS_P8_13_20", JavaFile(ref, P + "Filter8.java").floats("p8_13_20"), "ps/Filter8.java:17")
