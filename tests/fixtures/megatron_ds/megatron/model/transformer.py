# Excerpt-shaped fixture of megatron/model/transformer.py (anchor lines only).
FlashAttentionBuilder = get_accelerator().get_op_builder("FlashAttentionBuilder")
flash_attn_builder = None


class ParallelAttention(MegatronModule):
    def __init__(self, config, layer_number):
        args = get_args()
        self.enable_ds_sequence_parallel = parallel_state.get_sequence_parallel_world_size() > 1 \
                                           or args.force_ds_sequence_parallel
        if self.enable_ds_sequence_parallel:
            assert dist_attn_supported, 'Distributed attention is not supported in this DeepSpeed version'
            assert args.num_attention_heads % parallel_state.get_sequence_parallel_world_size() == 0
            self.dist_attn = DistributedAttention(local_attn, parallel_state.get_sequence_parallel_group())
        else:
            self.core_attention = local_attn
