"""``python -m bayesn_b200 input.yaml [overrides]`` - the reference's ``run_bayesn`` script (same arguments), and the
multi-GPU entry of ``run``: under ``torchrun --nproc-per-node N -m bayesn_b200 input.yaml`` every rank joins an NCCL
group, fits / trains on its block of SNe and rank 0 writes the outputs."""
import argparse
import os


def build_parser():
    p = argparse.ArgumentParser(prog='python -m bayesn_b200')
    p.add_argument('input', type=str)
    for name, typ in (('filters', str), ('outputdir', str), ('load_model', str), ('mode', str), ('num_chains', int),
                      ('fit_method', str), ('chain_method', str), ('initialisation', str), ('map', str), ('data_root', str),
                      ('data_table', str), ('version_photometry', str), ('num_warmup', int), ('num_samples', int),
                      ('snana', bool), ('outfile_prefix', str), ('sim_prescale', int), ('save_fit_errors', int),
                      ('error_floor', float)):
        p.add_argument(f'--{name}', type=typ, required=False)
    p.add_argument('--l_knots', type=float, required=False, nargs='*')
    p.add_argument('--tau_knots', type=float, required=False, nargs='*')
    p.add_argument('--drop_bands', type=str, required=False, nargs='*')
    p.add_argument('--jobsplit', type=int, nargs=2)
    p.add_argument('--private_data_path', type=str, required=False, nargs='*')
    return p


def main(argv=None):
    import yaml
    cmd_args = build_parser().parse_args(argv)
    if not os.path.exists(cmd_args.input):
        raise FileNotFoundError(f'Specified input file ({cmd_args.input}) was not found, please provide the path to an '
                                f'input yaml file or create an input.yaml in your current directory')
    with open(cmd_args.input, 'r') as fh:
        args = yaml.safe_load(fh)
    if 'load_model' not in args and cmd_args.load_model is None:
        args['load_model'] = 'T21_model'            # as the reference script
    elif cmd_args.load_model is not None:
        args['load_model'] = cmd_args.load_model
    if 'filters' not in args:
        args['filters'] = cmd_args.filters
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        if not dist.is_initialized():
            dist.init_process_group('nccl')
    from . import SEDmodel
    model = SEDmodel(load_model=args['load_model'], filter_yaml=args['filters'])
    try:
        return model.run(args, cmd_args)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == '__main__':
    main()
