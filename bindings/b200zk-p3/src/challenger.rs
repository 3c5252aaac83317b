//! Device-resident `DuplexChallenger<BabyBear, Poseidon2, 16, 8>`: lets the FRI commit phase (commit -> observe -> sample beta
//! -> fold, once per round) run without a host round trip per round.  The host challenger stays the source of truth for the
//! rest of the proof: `DeviceChallenger::from_host` replays its sponge state, `sync_back` writes the advanced state back.
use core::ptr;
use std::rc::Rc;

use b200zk_sys as sys;

use crate::ctx::{Ctx, Error};
use crate::F;

pub struct DeviceChallenger {
    pub(crate) ctx: Rc<Ctx>,
    pub(crate) raw: *mut sys::b200zk_chal,
}
impl DeviceChallenger {
    pub fn new(ctx: &Rc<Ctx>) -> Result<Self, Error> {
        let mut raw = ptr::null_mut();
        ctx.check(unsafe { sys::b200zk_chal_create(ctx.raw, &mut raw) })?;
        Ok(DeviceChallenger { ctx: ctx.clone(), raw })
    }
    pub fn observe(&mut self, values: &[F]) -> Result<(), Error> {
        self.ctx.check(unsafe { sys::b200zk_chal_observe(self.ctx.raw, self.raw, values.as_ptr() as *const u32, values.len() as u32) })
    }
    pub fn sample_vec(&mut self, n: usize) -> Result<Vec<F>, Error> {
        let mut out = F::zero_vec(n);
        self.ctx.check(unsafe { sys::b200zk_chal_sample(self.ctx.raw, self.raw, out.as_mut_ptr() as *mut u32, n as u32) })?;
        Ok(out)
    }
    pub fn sample_bits(&mut self, bits: usize) -> Result<usize, Error> {
        let mut out = 0u32;
        self.ctx.check(unsafe { sys::b200zk_chal_sample_bits(self.ctx.raw, self.raw, bits as u32, &mut out) })?;
        Ok(out as usize)
    }
    /// `GrindingChallenger::grind`: the smallest witness; canonical integer (p3 returns any valid one)
    pub fn grind(&mut self, bits: usize) -> Result<u32, Error> {
        let mut w = 0u32;
        self.ctx.check(unsafe { sys::b200zk_chal_grind(self.ctx.raw, self.raw, bits as u32, &mut w) })?;
        Ok(w)
    }
    /// sponge state (16) | input buffer (8) | fill | output buffer (8) | fill -- the fields of p3's DuplexChallenger
    pub fn state(&self) -> Result<[u32; 34], Error> {
        let mut s = [0u32; 34];
        self.ctx.check(unsafe { sys::b200zk_chal_state(self.ctx.raw, self.raw, s.as_mut_ptr()) })?;
        Ok(s)
    }
}
impl Drop for DeviceChallenger {
    fn drop(&mut self) {
        if !self.raw.is_null() {
            unsafe { sys::b200zk_chal_free(self.ctx.raw, self.raw) }
        }
    }
}
