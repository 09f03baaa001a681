"""b200sam — B200-native (sm_100a) implementation of the SAM pseudo-label refinement hot path of
multimodallearning/SamCarriesTheBurden.  Public surface mirrors the reference's:

    from samcarriestheburden_b200.segment_anything import sam_model_registry, SamPredictor
    from samcarriestheburden_b200.segment_anything.sam_mask_decoder_head import SAMMaskDecoderHead, EmbeddingStore
    from samcarriestheburden_b200.segment_anything.utils.prompt_utils import PromptExtractor, Prompt
    from samcarriestheburden_b200.utils.seg_refinement import SAMSegRefiner
"""
__version__ = "0.1.0"
