"""Stage files for the reference's YAML plugin point (see trajsde_b200/stages.py).  Import them from a process whose sys.path holds the
reference repository (train.py / test.py run from its root), e.g. in the config:

    encoder:
      file_path: <site>/trajsde_b200/plugins/enc_hivt_nusargo_sde_sep2_fused.py
      module_name: LocalEncoderSDESepPara2Fused
    decoder:
      file_path: <site>/trajsde_b200/plugins/dec_hivt_nusargo_sde_fused.py
      module_name: SDEDecoderFused
"""
