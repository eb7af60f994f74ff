"""Loss-landscape sweep around the propagation hot path -- mirror of the reference's VisTools/calc_loss.py
and the direction helpers of VisTools/directions.py that its notebook uses (experiment_vis_loss_landscape.ipynb cell 13).
Plotting (visualize.py, h52vtp.py) is out of scope."""
from .calc_loss import (calulate_single_element_loss_landscape, get_indices, overwrite_weights, read_surface_file,  # noqa: F401
                        setup_surface_file)
from .directions import create_random_direction, create_random_directions, get_weights  # noqa: F401
