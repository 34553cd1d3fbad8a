--[[
   nn.CageWarpVolumetric -- ALIGNet's warp sub-graph as ONE kernel (torch.CudaTensor only).

   Replaces, for a cage B x 3 x cD x cH x cW holding absolute (z, y, x) coordinates in [-1, 1] (the output of nn.Cumsum
   + beta, models/alignet_2d.lua:102-115), the two bilinearSampler() sub-graphs of models/alignet_2d.lua:136-147
   (models/alignet.lua:4-39) lifted to 3-D:

       field  = bilinearSampler()({cage, fieldI})      -- up-sample the cage to the output resolution
       output = bilinearSampler()({source, field})     -- warp the source

   updateOutput({source, cage})  ->  B x oD x oH x oW x C      (source B x D x H x W x C contiguous; output
                                                                resolution = depth / height / width, default the source's)
   updateGradInput({source, cage}, gradOutput)  ->  {gradSource, gradCage}

   The full-resolution field, the constant identity field `fieldI` and the field's gradient never exist in memory.
   Fields:
     self.onlyGrid   the source is data (nn.Identity, models/alignet_2d.lua:23-24): gradInput[1] is not computed (it is
                     returned as an EMPTY tensor, not 4 bytes per voxel of zeros)
     self.fastArith  VOLSTN_FAST_ARITH: the separable evaluation (brick kernels for contiguous single-channel
                     tensors); output within 1e-5 * max|output| of the bit-exact path (include/volstn_b200.h)
]]
local CageWarp, parent = torch.class('nn.CageWarpVolumetric', 'nn.Module')

-- flag bits of include/volstn_b200.h
local ONLY_GRID, ZERO_GRADVOL, FAST_ARITH = 1, 4, 0x40000

function CageWarp:__init(depth, height, width, onlyGrid, fastArith)
   parent.__init(self)
   self.depth, self.height, self.width = depth, height, width
   self.onlyGrid = onlyGrid or false
   self.fastArith = fastArith or false
   self.gradInput = {}
end

function CageWarp:check(input)
   local source, cage = input[1], input[2]
   assert(source:nDimension() == 5 and source:isContiguous(), 'source: contiguous B x D x H x W x C expected')
   assert(cage:nDimension() == 5 and cage:size(2) == 3 and cage:isContiguous(), 'cage: contiguous B x 3 x cD x cH x cW expected')
   assert(source:size(1) == cage:size(1))
end

function CageWarp:flags()
   local f = ZERO_GRADVOL
   if self.onlyGrid then f = f + ONLY_GRID end
   if self.fastArith then f = f + FAST_ARITH end
   return f
end

function CageWarp:updateOutput(input)
   self:check(input)
   local source, cage = input[1], input[2]
   if torch.type(self.output) ~= torch.type(source) then self.output = source.new() end
   self.output:resize(source:size(1), self.depth or source:size(2), self.height or source:size(3),
                      self.width or source:size(4), source:size(5))
   source.nn.CageWarpVolumetric_updateOutput(self, source, cage, self.output, self:flags())
   return self.output
end

function CageWarp:updateGradInput(input, gradOutput)
   self:check(input)
   local source, cage = input[1], input[2]
   self.gradInput[1] = self.gradInput[1] or source.new()
   if self.onlyGrid then
      self.gradInput[1]:resize(0)
   else
      self.gradInput[1]:resizeAs(source)
   end
   self.gradInput[2] = (self.gradInput[2] or source.new()):resizeAs(cage)
   source.nn.CageWarpVolumetric_updateGradInput(self, source, cage, self.gradInput[1], self.gradInput[2], gradOutput, self:flags())
   return self.gradInput
end
