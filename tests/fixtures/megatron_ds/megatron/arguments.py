# Excerpt-shaped fixture: the lines of Megatron-DeepSpeed's megatron/arguments.py that the integration patch anchors on.
import argparse


def _add_training_args(parser):
    group = parser.add_argument_group(title='training')
    group.add_argument('--sequence-parallel', action='store_true',
                       help='Enable Megatron-LM\'s sequence parallel optimization.')
    group.add_argument('--ds-sequence-parallel-size', type=int, default=1,
                       help='Enable DeepSpeed\'s sequence parallel. Cannot be combined with "--sequence-parallel", which enables Megatron-LM\'s sequence parallel.')
    group.add_argument('--force-ds-sequence-parallel', action='store_true',
                       help='use DeepSpeed sequence parallelism regardless of sequence parallel size.')
    return parser
