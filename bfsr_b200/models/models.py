"""Model registry with the reference's surface: `@register(name)`, `make(spec, args=None, load_sd=False)`.

Same contract as SRFlow-LP/code/models/models.py:7-23 and LINF-LP/models/models.py:7-23:
`spec = {'name': str, 'args': dict, 'sd': state_dict}` (the layout of the shipped `*-LP.pth` files,
LINF-LP/train.py:234-248).  The two reference projects register different classes under the name
'unet'; here one factory serves both and picks the variant from the arguments it is given.
"""
import copy

models = {}


def register(name):
    def decorator(cls):
        models[name] = cls
        return cls
    return decorator


def make(model_spec, args=None, load_sd=False):
    if args is not None:
        model_args = copy.deepcopy(model_spec['args'])
        model_args.update(args)
    else:
        model_args = model_spec['args']
    model = models[model_spec['name']](**model_args)
    if load_sd:
        model.load_state_dict(model_spec['sd'])
    return model
