"""Drop-in replacements for the reference's ``model`` package (model/stylegan2_common_layers.py,
model/stg2_generator.py, model/stg2_discriminator.py): same class names, constructor / forward signatures and
state_dict keys, arithmetic on the gif_b200 CUDA kernels."""
