--[[
   nn.Cumsum -- drop-in for Cumsum.lua (ALIGNet root; loaded by `require 'Cumsum'`, models/alignet_2d.lua:2,102).

   Input b x N x d1 x ... x dN: channel i is integrated (prefix sum) along spatial axis i; the gradient is the suffix
   sum along the same axes.  The reference does this with N x (split, squeeze, cumsum) and, backwards,
   N x (index, cumsum, index, copy) (Cumsum.lua:12-22, 39-48); here each direction is one native call.

   Difference on purpose: the reference's squeeze() (Cumsum.lua:18) also drops a batch of one, so it fails for b = 1;
   this module computes the documented semantics for every b.
]]
require 'volstn'

local Cumsum, parent = torch.class('nn.Cumsum', 'nn.Module')

function Cumsum:__init()
   parent.__init(self)
end

-- a contiguous version of t, buffered in self[field] when a copy is needed
local function packed(self, field, t)
   if t:isContiguous() then return t end
   self[field] = self[field] or t.new()
   self[field]:resizeAs(t):copy(t)
   return self[field]
end

function Cumsum:updateOutput(input)
   local x = packed(self, '_input', input)
   self.output:resizeAs(x)
   x.nn.Cumsum_updateOutput(self, x, self.output)
   return self.output
end

function Cumsum:updateGradInput(input, gradOutput)
   local g = packed(self, '_gradOutput', gradOutput)
   self.gradInput:resizeAs(input)
   g.nn.Cumsum_updateGradInput(self, g, self.gradInput)
   return self.gradInput
end

function Cumsum:clearState()
   nn.utils.clear(self, '_gradOutput', '_input')
   return parent.clearState(self)
end
