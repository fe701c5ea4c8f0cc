from .sp_transformer import SPBlock, SPTransformerConfig, SPTransformerLM, allreduce_sp_grads, rope

__all__ = ["SPBlock", "SPTransformerConfig", "SPTransformerLM", "allreduce_sp_grads", "rope"]
