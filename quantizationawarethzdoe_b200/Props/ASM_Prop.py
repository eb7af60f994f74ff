"""Band-limited angular-spectrum propagation -- drop-in for the reference's Props/ASM_Prop.py.

Same constructor and forward(field) -> ElectricField surface (Props/ASM_Prop.py:17-27, 314-378),
same attributes (`z` setter :185-195, `padding_scale`, `bandlimit_kernel`, `bandlimit_type`,
`do_padding`, `do_unpad_after_pad`, `check_Zc`, `shape`, buffers `_Kx/_Ky` :169-183), same
exceptions (:96, :309) -- but forward and backward run the fused sm_100a kernels behind
`thz_asm_propagate` (include/thzdoe.h): zero-pad, both FFTs, the transfer function H(kx,ky,lambda,z)
generated in registers, the crop, and (when the incoming field carries a deferred DOE modulation)
the DOE phase multiply and its adjoint, all without cuFFT / torch.fft.

`kernel_mode`:
  'auto' (default)       'inregister' when that stays within half of the 1e-5 parity budget for this geometry
                         (asm_host.inregister_deviation_estimate, evaluated once per plan: it does at the benchmark
                         physics, k z ~ 630 rad; it does not beyond z ~ 0.15 m at 1 mm wavelength), else 'cached'.
                         `resolved_kernel_mode` / `inregister_estimate` say what was picked and why.
  'inregister'           H is generated inside the column kernel from O(Hp+Wp) host-built vectors;
                         the band-limit mask is bit-identical to the reference, the phase differs from
                         the reference's only where torch's CPU sqrt is not correctly rounded.
  'cached'               H is built once per (shape, spacing, wavelengths, z) on the host with the
                         reference's own torch CPU ops, uploaded, and streamed by the kernel
                         (bit-identical H; +8 B per padded sample of HBM traffic per pass).
"""
import torch
import torch.nn as nn

from .. import asm_host as AH
from .. import bluestein as BL
from .. import functional as Fn
from .. import longline as LL
from ..DataType.ElectricField import DeferredElements, ElectricField


class ASM_prop(nn.Module):

    def __init__(self,
                 z_distance=0.0,
                 do_padding=True,
                 do_unpad_after_pad=True,
                 padding_scale=None,
                 bandlimit_kernel=True,
                 bandlimit_type='exact',
                 device=None,
                 kernel_mode='auto'):
        super().__init__()
        padding_scale = AH.normalise_padding_scale(padding_scale, do_padding)   # raises like ASM_Prop.py:96
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self._z = torch.tensor(z_distance, device=self.device)
        self.do_padding = do_padding
        self.do_unpad_after_pad = do_unpad_after_pad
        self.padding_scale = padding_scale
        self.bandlimit_kernel = bandlimit_kernel
        self.bandlimit_type = bandlimit_type
        if kernel_mode not in ('auto', 'inregister', 'cached'):
            raise ValueError("kernel_mode must be 'auto', 'inregister' or 'cached'")
        self.kernel_mode = kernel_mode
        self.resolved_kernel_mode = None if kernel_mode == 'auto' else kernel_mode
        self.inregister_estimate = None
        self._shape = None
        self.Kx = None
        self.Ky = None
        self.shape = None
        self.check_Zc = True
        self._plan_key = None
        self._plan = None

    # ---- reference-compatible helpers -------------------------------------------------------
    def compute_padding(self, H, W, return_size_of_padding=False):
        pad_h, pad_w, Hp, Wp = AH.compute_padding(int(H), int(W), self.padding_scale, self.do_padding)
        return (pad_h, pad_w) if return_size_of_padding else (Hp, Wp)

    def create_frequency_grid(self, H, W):
        with torch.no_grad():
            kx = (torch.linspace(0, H - 1, H) - (H // 2)) / H
            ky = (torch.linspace(0, W - 1, W) - (W // 2)) / W
            self.Kx, self.Ky = torch.meshgrid(kx, ky, indexing="ij")

    @property
    def shape(self):
        return self._shape

    @shape.setter
    def shape(self, shape):
        if shape is None:
            self._shape = None
            return
        old = self._shape
        self._shape = shape
        if old is None or old[-2] != shape[-2] or old[-1] != shape[-1]:
            self.create_frequency_grid(shape[-2], shape[-1])

    @property
    def Kx(self):
        return self._Kx

    @Kx.setter
    def Kx(self, Kx):
        self.register_buffer("_Kx", Kx)

    @property
    def Ky(self):
        return self._Ky

    @Ky.setter
    def Ky(self, Ky):
        self.register_buffer("_Ky", Ky)

    @property
    def z(self):
        return self._z

    @z.setter
    def z(self, z):
        # The reference stores whatever it is given (tensor, python float or numpy scalar; ASM_Prop.py:190-195).
        if isinstance(z, torch.Tensor) and z.device != self.device:
            z = z.to(self.device)
        self._z = z

    def _z_f32(self):
        """z as an fp32 host scalar.  Reading a CUDA tensor synchronises, so the value is cached per
        (tensor object, in-place version)."""
        z = self._z
        if isinstance(z, torch.Tensor):
            tok = (id(z), z._version)
            if getattr(self, "_z_tok", None) != tok:
                self._z_host = z.detach().to("cpu", torch.float32).reshape(())
                self._z_tok = tok
            return self._z_host
        return torch.tensor(float(z), dtype=torch.float32)

    def create_kernel(self, field):
        """Centred transfer function [1,C,Hp,Wp] complex64 exactly as the reference returns it
        (Props/ASM_Prop.py:212-311); host-built, for inspection / the cached mode."""
        H, W = field.shape[-2], field.shape[-1]
        Hp, Wp = self.compute_padding(H, W)
        self.shape = torch.Size((field.shape[0], field.shape[1], Hp, Wp))
        Hc = AH.tf_centred_reference_order(Hp, Wp, field.spacing, field.wavelengths, self._z_f32(),
                                           self.bandlimit_kernel, self.bandlimit_type)
        return Hc[None].to(self.device)

    # ---- plan cache ------------------------------------------------------------------------
    def _get_plan(self, B, C, H, W, spacing, wavelengths, device):
        z = self._z_f32()
        # fast path: same tensor objects (unchanged in place) as last time -> same plan, no device->host reads
        fast = (id(spacing), spacing._version, id(wavelengths), wavelengths._version, C, H, W, float(z), str(device),
                self.do_padding, self.do_unpad_after_pad, self.bandlimit_kernel, self.bandlimit_type, self.kernel_mode,
                id(self.padding_scale))
        if self._plan is not None and getattr(self, "_fast_key", None) == fast:
            self._plan.B = B
            return self._plan
        key = (C, H, W, tuple(spacing.detach().cpu().reshape(-1).tolist()),
               tuple(wavelengths.detach().cpu().reshape(-1).tolist()), float(z), str(device),
               self.do_padding, self.do_unpad_after_pad,
               None if self.padding_scale is None else tuple(self.padding_scale.reshape(-1).tolist()),
               self.bandlimit_kernel, self.bandlimit_type, self.kernel_mode)
        if key != self._plan_key:
            if self.bandlimit_kernel and self.bandlimit_type not in ('exact', 'approx'):
                raise Exception("Should not be in this state.")                      # ASM_Prop.py:309
            pad_h, pad_w, Hp, Wp = AH.compute_padding(H, W, self.padding_scale, self.do_padding)
            self.shape = torch.Size((B, C, Hp, Wp))
            if self.bandlimit_kernel and self.check_Zc is True:                       # ASM_Prop.py:279-285
                Zc = AH.critical_distance(Hp, spacing, wavelengths)
                if float(z) > float(Zc):
                    print("The propagation distance is greater than critical distance {} m, the TF will be undersampled!".format(Zc.numpy()))
                else:
                    print("The critical distance is {} m, the TF will be fine during the sampling !".format(Zc.numpy()))
                self.check_Zc = False
            unpad = bool(self.do_padding and self.do_unpad_after_pad)
            sp = None
            if LL.needs_split(Hp, Wp):
                try:
                    sp = LL.split_or_none(Hp, Wp)
                except NotImplementedError:
                    sp = None                         # not 2 or 4 times a radix-plan length: chirp-z below (or its size error)
            if sp is not None:
                # an edge above 16384 points: one outer decimation step per long axis around the fused pipeline (longline.py);
                # the transfer function of every sub-problem is the decimated one, same two modes as below
                rowvec, colvec, scal = AH.tf_vectors(Hp, Wp, spacing, wavelengths, z, self.bandlimit_kernel, self.bandlimit_type)
                mode_ = self.kernel_mode
                if mode_ == 'auto':
                    self.inregister_estimate = AH.inregister_estimate_for(Hp, Wp, spacing, wavelengths, z, self.bandlimit_kernel,
                                                                          self.bandlimit_type)
                    mode_ = 'inregister' if self.inregister_estimate <= AH.INREGISTER_BUDGET else 'cached'
                sv = LL.split_tf_vectors(rowvec, colvec, scal, *sp) if mode_ == 'inregister' else None
                if sv is not None:
                    self._plan = LL.SplitAsmPlan(B, C, H, W, pad_h, pad_w, Hp, Wp, unpad, device, sp[0], sp[1], vectors=sv)
                elif device.type == "cuda":
                    table = LL.split_tf_table_device(Hp, Wp, spacing, wavelengths, z, self.bandlimit_kernel, self.bandlimit_type,
                                                     device, *sp)
                    self._plan = LL.SplitAsmPlan(B, C, H, W, pad_h, pad_w, Hp, Wp, unpad, device, sp[0], sp[1], table=table)
                else:
                    Hc = AH.tf_centred_reference_order(Hp, Wp, spacing, wavelengths, z, self.bandlimit_kernel, self.bandlimit_type)
                    table = LL.split_tables_from_natural(torch.fft.ifftshift(Hc, dim=(-2, -1)), *sp)
                    self._plan = LL.SplitAsmPlan(B, C, H, W, pad_h, pad_w, Hp, Wp, unpad, device, sp[0], sp[1], table=table)
                self.resolved_kernel_mode = '%s (split %d x %d)' % ('inregister' if sv is not None else 'cached', sp[0], sp[1])
                self._plan_key = key
                self._fast_key = fast
                self._fast_refs = (spacing, wavelengths)
                return self._plan
            if not (BL.length_supported(Hp) and BL.length_supported(Wp)):
                # an edge length with a prime factor > 7 (the reference's torch.fft takes any size): chirp-z on the fused
                # pipeline, transfer function = the reference's own (host-built, as in 'cached')
                Hc = AH.tf_centred_reference_order(Hp, Wp, spacing, wavelengths, z, self.bandlimit_kernel, self.bandlimit_type)
                unpad = bool(self.do_padding and self.do_unpad_after_pad)
                self._plan = BL.BluesteinAsmPlan(C, H, W, pad_h, pad_w, Hp, Wp, unpad, device, torch.fft.ifftshift(Hc, dim=(-2, -1)))
                self.resolved_kernel_mode = 'cached (chirp-z, length %d x %d)' % (Hp, Wp)
                self._plan_key = key
                self._fast_key = fast
                self._fast_refs = (spacing, wavelengths)
                return self._plan
            rowvec, colvec, scal = AH.tf_vectors(Hp, Wp, spacing, wavelengths, z, self.bandlimit_kernel, self.bandlimit_type)
            table, mode = None, 0
            chunked = AH.row_vectors_chunked(Hp)
            mode_ = self.kernel_mode
            if mode_ == 'auto':
                self.inregister_estimate = AH.inregister_estimate_for(Hp, Wp, spacing, wavelengths, z, self.bandlimit_kernel,
                                                                      self.bandlimit_type)
                mode_ = 'inregister' if self.inregister_estimate <= AH.INREGISTER_BUDGET else 'cached'
            dv = AH.tf_device_vectors(rowvec, colvec, scal, chunked=chunked) if mode_ == 'inregister' else None
            if dv is not None:
                rowvec, colvec, scal = dv
            self.resolved_kernel_mode = 'inregister' if dv is not None else 'cached'
            if mode_ == 'cached' or dv is None:
                # the reference's own angles (host, unique quarter) expanded on the device; whole table on the host only if the
                # mask cannot be folded into thresholds
                table = AH.tf_table_device(Hp, Wp, spacing, wavelengths, z, self.bandlimit_kernel, self.bandlimit_type, device) \
                    if device.type == "cuda" else None
                if table is None:
                    Hc = AH.tf_centred_reference_order(Hp, Wp, spacing, wavelengths, z, self.bandlimit_kernel, self.bandlimit_type)
                    table = AH.tf_table_slot_order(Hc)
                mode = 1
            unpad = bool(self.do_padding and self.do_unpad_after_pad)
            self._plan = Fn.AsmPlan(B, C, H, W, pad_h, pad_w, Hp, Wp, unpad, device, rowvec, colvec, scal, table, mode,
                                    row_chunked=chunked)
            self._plan_key = key
        self._fast_key = fast
        self._fast_refs = (spacing, wavelengths)      # keep the keyed tensors alive so their ids stay unique
        self._plan.B = B
        return self._plan

    # ---- forward ---------------------------------------------------------------------------
    def forward(self, field):
        wavelengths = field.wavelengths
        deferred = getattr(field, "_deferred", None)
        if deferred is not None and getattr(deferred, "is_elements", False):
            deferred = DeferredElements.pending(field)          # counts the use; a re-used chain is evaluated once instead
        if deferred is not None and getattr(field, "_data", None) is None:
            B, C, H, W = deferred.shape
            dev = deferred.device
        else:
            deferred = None
            data = field.data
            B, C, H, W = data.shape
            dev = data.device
        plan = self._get_plan(B, C, H, W, field.spacing, wavelengths, dev)
        if isinstance(plan, BL.BluesteinAsmPlan):      # DOE modulation (if any) is materialised by its own kernel first
            out = BL.BluesteinAsmFn.apply(field.data, plan)
            return ElectricField(data=out, wavelengths=wavelengths, spacing=field.spacing, device=dev)
        if isinstance(plan, LL.SplitAsmPlan):          # likewise on canvases above 16384 points per edge
            out = Fn.AsmPropagateFn.apply(field.data, plan)
            return ElectricField(data=out, wavelengths=wavelengths, spacing=field.spacing, device=dev)
        if deferred is not None and getattr(deferred, "is_elements", False):        # aperture / lens only: fused on load
            out = Fn.AsmPropagateFn.apply(deferred.x, plan, deferred.mask, deferred.mul)
        elif deferred is not None:                                                  # [aperture / lens +] DOE: fused on load
            out = Fn.DoeAsmFn.apply(deferred.x, deferred.height_map, plan, deferred.coef, deferred.mask, deferred.mul,
                                    getattr(deferred, "reducer", None), getattr(deferred, "levels", None))
        else:
            out = Fn.AsmPropagateFn.apply(data, plan)
        return ElectricField(data=out, wavelengths=wavelengths, spacing=field.spacing, device=dev)
