"""quantizationawarethzdoe_b200 -- B200-native (sm_100a) field-propagation hot path of
sihan-shao/QuantizationAwareTHzDOE behind the reference's own torch.nn.Module surface.

    from quantizationawarethzdoe_b200 import ElectricField, ASM_prop, CZT_prop, STEQuantizedDOELayer

The sub-packages mirror the reference layout (DataType/, Props/, Components/, utils/), so
`from quantizationawarethzdoe_b200.Props.ASM_Prop import ASM_prop` is the one-line switch.
All arithmetic runs in csrc/libthzdoe.so (C ABI: include/thzdoe.h); there is no CPU fallback.
"""
from .DataType.ElectricField import ElectricField
from .Props.ASM_Prop import ASM_prop
from .Components.QuantizedDOE import (
    DOELayer, FixDOEElement, FullPrecisionDOELayer, STEQuantizedDOELayer, PSQuantizedDOELayer,
    SoftGumbelQuantizedDOELayer, SoftGumbelQuantizedDOELayerv2, SoftGumbelQuantizedDOELayerv3,
    NaiveGumbelQuantizedDOELayer, STEQuantizationFunction, ste_quan,
)

from .Props.CZT_Prop import CZT_prop
from .Props.RSC_Prop import RSC_prop, VRS_prop
from .Components.Thin_Lens import Thin_LensElement
from .Components.Aperture import ApertureElement
from .LightSource.Gaussian_beam import Guassian_beam
from .train import FusedAdam, normalized_intensity_mse

__version__ = "0.1.0"
