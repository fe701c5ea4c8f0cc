# Excerpt-shaped fixture of megatron/initialize.py (anchor lines only).
def _initialize_distributed():
    args = get_args()
    if True:
        if True:
            mpu.initialize_model_parallel(args.tensor_model_parallel_size,
                                           args.pipeline_model_parallel_size,
                                           args.ds_sequence_parallel_size,
                                           args.virtual_pipeline_model_parallel_size,
                                           args.pipeline_model_parallel_split_rank,
                                           use_distributed_optimizer=args.use_distributed_optimizer)
            if args.rank == 0:
                print('> initialized')
