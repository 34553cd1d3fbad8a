-- volstn (B200): package loader, same shape as stn3dtorch/init.lua:1-12.
-- `libvolstn` / `libcuvolstn` are built from lua/init.c / lua/init.cu and link libvolstn_b200.so.
require 'nn'
local withCuda = pcall(require, 'cutorch')

require 'libvolstn'
if withCuda then
   require 'libcuvolstn'
end

require('volstn.VolumetricGridGenerator')
require('volstn.BilinearSamplerVolumetric')
require('volstn.Cumsum')
require('volstn.AffineSamplerVolumetric')           -- grid generator + sampler in one call (FloatTensor and CudaTensor)
if withCuda then
   require('volstn.CageWarpVolumetric')            -- fused modules (CudaTensor only; INTEGRATION.md section 3)
   require('volstn.CageWarpSmoothL1Volumetric')
   require('volstn.BilinearSamplerSmoothL1Volumetric')
end

return nn
