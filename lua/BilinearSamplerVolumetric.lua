--[[
   nn.BilinearSamplerVolumetric -- drop-in for stn3dtorch/BilinearSamplerVolumetric.lua over libvolstn_b200.

   Same class name, constructor, updateOutput({inputImages, grids}) / updateGradInput({inputImages, grids}, gradOutput),
   same layouts (inputImages B x D x H x W x C contiguous, grids B x D x H x W x 4 holding (z, y, x, .) in [-1, 1],
   zero padding outside), same 4-D convenience form, same serialised fields (output, gradInput).

   What changed underneath: the native call goes to the B200 kernels, and the zero-fill of gradInput[1] travels with
   it as a flag in the optional trailing number -- the slot the reference's C code reads as `focal_length` and ignores
   (generic/BilinearSamplerVolumetric.c:447,592) -- instead of being a separate pass here; gradInput[2] needs no
   zero-fill at all because every element of it is assigned (...c:819-822).

   Optional fields (nil in models serialised by the reference = the reference's behaviour):
     self.onlyGrid    skip gradInput[1] (the `onlyGrid` switch the reference declares and never enables, ...c:594)
     self.onlyVolume  skip gradInput[2]
]]
local Sampler, parent = torch.class('nn.BilinearSamplerVolumetric', 'nn.Module')

-- flag bits of include/volstn_b200.h
local ONLY_GRID, ONLY_VOLUME, ZERO_GRADVOL = 1, 2, 4

function Sampler:__init()
   parent.__init(self)
   self.gradInput = {}
end

function Sampler:check(input, gradOutput)
   local images, grids = input[1], input[2]
   assert(images:isContiguous(), 'Input images have to be contiguous')
   assert(images:nDimension() == 5)
   assert(grids:nDimension() == 5)
   assert(images:size(1) == grids:size(1))   -- batch
   assert(grids:size(5) == 4)                -- (z, y, x, .)
   if gradOutput then
      for d = 1, 4 do                        -- batch, depth, height, width follow the grid
         assert(grids:size(d) == gradOutput:size(d))
      end
   end
end

-- a 4-D sample as a batch of one (a view, no copy)
local function batched(t)
   if t:nDimension() ~= 4 then return t end
   return t:view(1, t:size(1), t:size(2), t:size(3), t:size(4))
end

function Sampler:updateOutput(input)
   local single = input[1]:nDimension() == 4
   local images, grids = batched(input[1]), batched(input[2])
   self:check({images, grids})
   self.output:resize(images:size(1), grids:size(2), grids:size(3), grids:size(4), images:size(5))
   images.nn.BilinearSamplerVolumetric_updateOutput(self, images, grids, self.output)
   if single then
      self.output = self.output:select(1, 1)
   end
   return self.output
end

function Sampler:updateGradInput(input, gradOutput)
   local single = input[1]:nDimension() == 4
   local images, grids, gradOut = batched(input[1]), batched(input[2]), batched(gradOutput)
   self:check({images, grids}, gradOut)
   self.gradInput[1] = (self.gradInput[1] or images.new()):resizeAs(images)
   self.gradInput[2] = (self.gradInput[2] or images.new()):resizeAs(grids)
   local flags = ZERO_GRADVOL
   if self.onlyGrid then
      flags = flags + ONLY_GRID
      self.gradInput[1]:zero()
   elseif self.onlyVolume then
      flags = flags + ONLY_VOLUME
      self.gradInput[2]:zero()
   end
   images.nn.BilinearSamplerVolumetric_updateGradInput(self, images, grids, self.gradInput[1], self.gradInput[2], gradOut, flags)
   if single then
      self.gradInput[1] = self.gradInput[1]:select(1, 1)
      self.gradInput[2] = self.gradInput[2]:select(1, 1)
   end
   return self.gradInput
end
