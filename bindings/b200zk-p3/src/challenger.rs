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
    /// Take over a host transcript: p3's `DuplexChallenger` keeps `sponge_state`, `input_buffer` and `output_buffer` public, and
    /// the device challenger has exactly those fields (`b200zk_chal_set_state`).
    pub fn from_host<P>(ctx: &Rc<Ctx>, host: &p3_challenger::DuplexChallenger<F, P, 16, 8>) -> Result<Self, Error>
    where
        P: p3_symmetric::CryptographicPermutation<[F; 16]>,
    {
        let c = Self::new(ctx)?;
        let mut s = [0u32; 34];
        for (d, v) in s[..16].iter_mut().zip(host.sponge_state.iter()) {
            *d = crate::monty_bits(*v);
        }
        for (d, v) in s[16..24].iter_mut().zip(host.input_buffer.iter()) {
            *d = crate::monty_bits(*v);
        }
        s[24] = host.input_buffer.len() as u32;
        // p3 pops sampled values from the END of output_buffer; the device keeps the same vector and fill count
        for (d, v) in s[25..33].iter_mut().zip(host.output_buffer.iter()) {
            *d = crate::monty_bits(*v);
        }
        s[33] = host.output_buffer.len() as u32;
        ctx.check(unsafe { sys::b200zk_chal_set_state(ctx.raw, c.raw, s.as_ptr()) })?;
        Ok(c)
    }
    /// Hand the advanced transcript back to the host challenger (after the FRI commit phase, PoW and query sampling).
    pub fn sync_back<P>(&self, host: &mut p3_challenger::DuplexChallenger<F, P, 16, 8>) -> Result<(), Error>
    where
        P: p3_symmetric::CryptographicPermutation<[F; 16]>,
    {
        let s = self.state()?;
        let f = |w: u32| -> F { unsafe { core::mem::transmute::<u32, F>(w) } };   // Montgomery bits, repr(transparent)
        for (d, w) in host.sponge_state.iter_mut().zip(&s[..16]) {
            *d = f(*w);
        }
        host.input_buffer = s[16..16 + s[24] as usize].iter().map(|w| f(*w)).collect();
        host.output_buffer = s[25..25 + s[33] as usize].iter().map(|w| f(*w)).collect();
        Ok(())
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
