// The ragged walk, second generation ("chain walk"): included by lines8.cuh, shared by lines8_kernel and linesq_kernel.
//
// Lines of any length, one per lane, as in l8_run_ragged - but
//   * a lane is not bound to a line of a tile pair: lines are handed out in order to whichever lane is free, and the two
//     tile buffers of the warp ROLL: as soon as every line of a tile is finished the buffer is refilled with the next tile
//     while the lanes work on the other one.  No sorting, no pairing, and the copies overlap the walk.
//   * find()'s table-driven reverse pass (indexBackwards, DFAClassBuilder.java:529-586) is CHAINED onto the forward walk
//     of the same lane as a second phase of the same loop: the window is read at a descending address and byte-reversed
//     (one PRMT per word with a per-lane selector), after which walking it forwards through the BACKWARDS rows of the
//     table image is exactly the backward walk.  One loop body for both directions, so lanes in different phases do not
//     diverge - the separate pooled reverse pass of the first generation ran at 26 % lane utilisation and was 28 % of all
//     instructions on the e-mail regex.
//   * lines that cannot match are never walked: when every accepting path of the automaton takes a transition on one
//     class of chars (the `@` of the e-mail regex - what the reference's Factorization calls a required factor and seeks
//     with indexOf, DFAClassBuilder.java:394-399), a tile is first scanned for that class with packed compares, 16 bytes
//     per lane and step, and only the lines whose 16-byte chunks contain such a char are handed out (a neighbour's char in
//     a shared chunk makes a line survive needlessly - never the other way round); the rest get "no match" directly.
// Results are identical to the generated loops of the reference (indexForwards :335-471, indexBackwards :529-614, glue
// :616-667): per line, the forward phase is l8_run_ragged's walk and the reverse phase is l8_reverse's.
#pragma once

namespace ndl {

template <int CM>
__device__ __forceinline__ void l8_run_chain(const Lines8Params& p, const L8Ctx& cx, const uint32_t buf0, const uint32_t buf1,
                                             const uint32_t lane, const uint32_t warp_global, const uint32_t n_warps) {
  using CharT = typename std::conditional<L8Chars<CM>::kBytes == 1, uint8_t, uint16_t>::type;
  constexpr uint32_t kCharBytes = L8Chars<CM>::kBytes;
  constexpr uint32_t kPer = L8Chars<CM>::kPerChunk;
  constexpr uint32_t kFull = 0xffffffffu;
  constexpr uint32_t kStateMask = L8Enc<CM>::kStateMask;
  const BatchParams& g = p.g;
  const uint8_t* const data = static_cast<const uint8_t*>(g.data);
  const uint32_t n = static_cast<uint32_t>(g.n);
  const uint32_t per_warp = (n + n_warps - 1) / n_warps;
  const uint32_t lo = min(n, warp_global * per_warp), hi = min(n, lo + per_warp);
  constexpr uint32_t kCap = kL8WarpBuf - 16;  // the last 16 bytes stay free for the window that runs past the tile
  const uint32_t mode = static_cast<uint32_t>(g.mode);
  const bool use_from = g.from != nullptr && mode == 2;
  // the reverse pass runs on the staged tile, chained onto the forward walk
  const bool chain_rev = mode == 2 && g.reverse_mode == 0 && p.has_bwd != 0;
  const bool filter = kCharBytes == 1 && p.flt_on != 0;
  const uint32_t rev_sel = kCharBytes == 1 ? 0x4567u : 0x5476u;  // PRMT selector that reverses the chars of a word
  const int32_t w_root = g.fwd.root_accepting ? 0 : -1;
  const uint32_t tail_root = g.fwd.root_accepting ? 1u : 0u;
  const uint32_t lt_mask = (1u << lane) - 1u;

  // ---- tiles.  A tile is the longest run of <= 32 consecutive lines whose bytes (from the 16-byte boundary below the
  // first line) fit a buffer.  Line descriptor: first byte relative to the buffer | length in chars << 12 | line << 23.
  uint32_t next_line = lo;                  // first line that has not been staged
  uint32_t c0 = 0, c1 = 0;                  // first line of the tile in buffer 0 / 1
  uint32_t n0 = 0, n1 = 0;                  // lines of it to walk (after the filter), fin: how many of them are done
  uint32_t fin0 = 0, fin1 = 0;
  uint32_t tdesc0 = 0, tdesc1 = 0;          // per lane: descriptor of line `lane` of the staged tile
  int32_t tfrm0 = 0, tfrm1 = 0;             // per lane: its find(from, to) start offset
  uint32_t flags = 0;                       // bit b: buffer b holds a staged tile no line of which has been handed out;
                                            // bit 2: the next lines are long ones, to be streamed once both buffers are idle
  // the tile lines are being handed out from
  uint32_t d_buf = 0, d_n = 0, drawn = 0, d_c = 0, d_base = buf0;
  uint32_t d_desc = 0;                      // per lane: descriptor of the lane-th line to hand out
  int32_t d_frm = 0;

  auto try_stage = [&](uint32_t b) {  // buffer b is free: bring in the next tile
    const uint32_t c = next_line, buf = b ? buf1 : buf0;
    const uint64_t s0 = batch_off(g, c) * kCharBytes;  // bytes
    const uint32_t idx = min(c + lane + 1, hi);
    const uint64_t e_off = batch_off(g, idx) * kCharBytes;
    const uint32_t slack = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(data) + s0) & 15u;
    const uint64_t rel_end = e_off - s0 + slack;  // end of this lane's line relative to the tile buffer
    const bool fits = (c + lane < hi) && rel_end <= kCap;
    const uint32_t count = __popc(__ballot_sync(kFull, fits));  // offsets are non-decreasing, so `fits` is a prefix
    if (count < 8 && count < hi - c) {  // long lines: nothing staged, streamed later
      flags |= 4u;
      return;
    }
    const uint32_t end32 = static_cast<uint32_t>(rel_end);
    uint32_t prev = __shfl_up_sync(kFull, end32, 1);
    if (lane == 0) prev = slack;
    const uint32_t desc = prev | ((end32 - prev) / kCharBytes) << 12 | lane << 23;
    int32_t frm = 0;
    if (use_from && lane < count) frm = g.from[c + lane];
    const uint32_t total_bytes = __shfl_sync(kFull, end32, count - 1);
    const uint32_t n_chunks = (total_bytes + 15) >> 4;
    const uint8_t* src = data + s0 - slack;
    for (uint32_t j = lane; j < n_chunks; j += 32) cp_async16(buf + l8_rslot(j), src + (static_cast<uint64_t>(j) << 4));
    cp_async_commit();
    if (b == 0) { c0 = c; n0 = count; fin0 = 0; tdesc0 = desc; tfrm0 = frm; }
    else        { c1 = c; n1 = count; fin1 = 0; tdesc1 = desc; tfrm1 = frm; }
    flags |= 1u << b;
    next_line = c + count;
  };
  auto restage = [&]() {  // refill every buffer whose tile is finished
    if (next_line < hi && !(flags & 4u)) {
      if (!(flags & 1u) && fin0 == n0) {
        __syncwarp();
        try_stage(0);
      }
      if (next_line < hi && !(flags & 6u) && fin1 == n1) {
        __syncwarp();
        try_stage(1);
      }
    }
  };
  auto start_tile = [&](uint32_t b) {  // begin handing out the staged tile of buffer b: its copies must have landed
    cp_async_wait<0>();
    __syncwarp();
    flags &= ~(1u << b);
    d_buf = b;
    d_base = b ? buf1 : buf0;
    d_c = b ? c1 : c0;
    drawn = 0;
    uint32_t count = b ? n1 : n0;
    uint32_t desc = b ? tdesc1 : tdesc0;
    int32_t frm = b ? tfrm1 : tfrm0;
    if (filter) {
      // which 16-byte chunks of the tile hold a char of the required class?  Lane l looks at chunks l, l + 32, l + 64, l + 96.
      uint32_t h[4];
#pragma unroll
      for (uint32_t k = 0; k < 4; k++) {
        const uint4 v = lds_data16(d_base + l8_rslot(lane + 32 * k));
        uint32_t any = 0;
        const uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const uint32_t w80 = words[j] | 0x80808080u;
          any |= ((w80 - p.flt_lo) ^ (w80 - p.flt_hi)) & ~words[j];
        }
        h[k] = __ballot_sync(kFull, (any & 0x80808080u) != 0);
      }
      const uint32_t start = desc & 0xfffu, len = (desc >> 12) & 0x7ffu;
      const uint32_t ch_lo = start >> 4, ch_hi = (start + len + 15) >> 4;  // the line's chunks: [ch_lo, ch_hi)
      const uint32_t cnt = ch_hi - ch_lo, k = ch_lo >> 5;
      const uint32_t w_lo = k == 0 ? h[0] : k == 1 ? h[1] : k == 2 ? h[2] : h[3];
      const uint32_t w_hi = k == 0 ? h[1] : k == 1 ? h[2] : k == 2 ? h[3] : 0u;
      const uint32_t bits = __funnelshift_r(w_lo, w_hi, ch_lo & 31u);  // chunk ch_lo is bit 0
      // (a negative `from` is outside the reference's domain: such a line goes to the generic walk as before)
      const bool survive = lane < count && ((len != 0 && (cnt > 32 || (bits & (cnt >= 32 ? kFull : (1u << cnt) - 1u)) != 0)) ||
                                            (use_from && frm < 0));
      if (lane < count && !survive) {  // no char of the required class: no match
        const uint64_t i = static_cast<uint64_t>(d_c) + lane;
        g.matched[i] = 0;
        if (mode == 2) {
          g.start[i] = -1;
          g.end[i] = -1;
        }
      }
      const uint32_t sv = __ballot_sync(kFull, survive);
      count = __popc(sv);
      const uint32_t src = __fns(sv, 0, static_cast<int>(lane) + 1) & 31u;  // lane-th surviving line
      desc = __shfl_sync(kFull, desc, src);
      if (use_from) frm = __shfl_sync(kFull, frm, src);
      if (b) n1 = count; else n0 = count;
    }
    d_n = count;
    d_desc = desc;
    d_frm = frm;
  };

  // ---- the lane's current job
  uint32_t active = 0, rev = 0, jslow = 0;
  uint32_t jbase = 0, jline = 0, jps = 0, jlen = 0, jtile = 0;
  int32_t jfrom = 0, last_f = -1;
  uint32_t total = 0, pos = 0;  // chars of this phase, and how many of them have been walked
  int32_t win = 0;              // byte address, relative to the buffer, of the next 16-byte window (may be < 0 going backwards)
  int32_t win_step = 16;
  uint32_t e = 0, dead = 0;
  int32_t w = -1;               // chars walked up to and including the last accepting step of this phase, or -1
  uint32_t tail_bit = 0;

  if (next_line < hi) try_stage(0);
  if (next_line < hi && !(flags & 4u)) try_stage(1);

  for (;;) {
    // ---- hand out lines to the lanes that have none
    const uint32_t need = __ballot_sync(kFull, active == 0);
    if (need) {
      if (drawn >= d_n) {  // the tile being handed out is exhausted: refill what is free, move on to a staged tile
        restage();
        if (flags & (1u << (d_buf ^ 1u))) start_tile(d_buf ^ 1u);
        else if (flags & (1u << d_buf)) start_tile(d_buf);
      }
      if (drawn < d_n) {
        const uint32_t j = drawn + __popc(need & lt_mask);
        const uint32_t d = __shfl_sync(kFull, d_desc, j & 31u);
        int32_t f = 0;
        if (use_from) f = __shfl_sync(kFull, d_frm, j & 31u);
        if (active == 0 && j < d_n) {
          const uint32_t start = d & 0xfffu, len = (d >> 12) & 0x7ffu;
          // from outside [0, len) keeps the reference's corner cases: the generic walk, at the end of this (empty) job
          jslow = (f < 0 || (f != 0 && static_cast<uint32_t>(f) >= len)) ? 1u : 0u;
          jfrom = jslow ? 0 : f;
          jline = d_c + (d >> 23);
          jtile = d_buf;
          jbase = d_base;
          jps = start + static_cast<uint32_t>(jfrom) * kCharBytes;
          jlen = len - static_cast<uint32_t>(jfrom);
          active = 1;
          rev = 0;
          total = jslow ? 0u : jlen;
          pos = 0;
          win = static_cast<int32_t>(jps);
          win_step = 16;
          e = cx.root;
          dead = cx.fwd_dead;
          w = w_root;
          tail_bit = tail_root;
        }
        drawn += __popc(need);
      } else if (need == kFull) {
        // nothing is being walked and nothing can be handed out: lines left to start, to stream, to stage?
        if (flags & 3u) continue;  // (a tile whose lines were all filtered out has just been replaced)
        if (flags & 4u) {          // long lines: stream the next 32, one per lane, through both (idle) buffers
          cp_async_wait<0>();
          __syncwarp();
          const uint32_t m = min(32u, hi - next_line);
          l8_stream_group<CM, CharT>(p, cx, buf0, buf1, lane, next_line, m, use_from);
          next_line += m;
          flags &= ~4u;
          continue;
        }
        if (next_line < hi) continue;  // both buffers are free: the next pass stages
        break;
      }
    }

    // ---- one 16-byte window of the lane's line, in walk order
    uint32_t finished = 0;
    if (active) {
      if (pos < total) {
        const int32_t q = win >> 4;
        const uint4 x = lds_data16(jbase + l8_rslot(static_cast<uint32_t>(q > 0 ? q : 0)));
        const uint4 y = lds_data16(jbase + l8_rslot(static_cast<uint32_t>(q + 1)));
        uint4 wv = L8Align(static_cast<uint32_t>(win) & 15u).apply(x, y);
        if (chain_rev) {  // going backwards: the chars of the window in reverse order
          const uint32_t sel = rev ? rev_sel : 0x3210u;
          wv = make_uint4(__byte_perm(wv.x, wv.w, sel), __byte_perm(wv.y, wv.z, sel), __byte_perm(wv.z, wv.y, sel),
                          __byte_perm(wv.w, wv.x, sel));
        }
        uint32_t mask = 0;
        l8_chunk<CM>(wv, p.q, cx, e, mask);
        const uint32_t valid = min(kPer, total - pos);
        mask >>= (kPer - valid);  // drop the accept bits of the chars past the end of the phase
        const int32_t cand = static_cast<int32_t>(pos + valid + 1) - __ffs(mask);
        w = mask ? cand : w;
        tail_bit = mask & 1u;
        pos += kPer;
        win += win_step;
        // a dead automaton stays dead (and never accepts); containedIn() has its answer at the first accept
        if ((e & kStateMask) == dead || (mode == 1 && w != -1)) pos = total;
      }
      if (pos >= total) {
        if (chain_rev && !rev && w != -1) {
          // found the end of the match: indexBackwards(end - 1, from) over the same staged line, second phase
          last_f = w;
          rev = 1;
          total = static_cast<uint32_t>(w);
          pos = 0;
          win = static_cast<int32_t>(jps + total * kCharBytes) - 16;
          win_step = -16;
          e = cx.bwd_root;
          dead = cx.bwd_dead;
          w = g.bwd.root_accepting ? static_cast<int32_t>(total) : -1;  // lastMatch = lowerBound when the root accepts (:543-547)
        } else {
          if (jslow) {
            l8_slow_line<CharT>(g, jline);
          } else if (mode == 2 && (rev || w == -1 || g.reverse_mode == 2)) {
            // find(): the glue of DFAClassBuilder.java:625-659 for the cases settled here
            int32_t last = rev ? last_f : w, st = -1;
            if (last != -1) {
              if (rev) st = w == -1 ? 0x7fffffff : static_cast<int32_t>(total) - w + jfrom;
              else st = last + jfrom - g.min_length;  // start = end - minLength (:640-646)
              last += jfrom;
            }
            g.matched[jline] = last != -1;
            g.start[jline] = st;
            g.end[jline] = last;
          } else {
            const uint32_t base = jbase;
            l8_finish<CM, CharT>(p, cx, jline, jlen, w, tail_bit != 0, jps, [&](uint32_t ch) { return base + l8_rslot(ch); }, jfrom);
          }
          active = 0;
          finished = 1;
        }
      }
    }
    const uint32_t f0 = __ballot_sync(kFull, finished != 0 && jtile == 0);
    const uint32_t f1 = __ballot_sync(kFull, finished != 0 && jtile != 0);
    if (f0 | f1) {
      fin0 += __popc(f0);
      fin1 += __popc(f1);
      restage();  // a finished tile's buffer is refilled while the lanes work on the other one
    }
  }
  cp_async_wait<0>();
}

}  // namespace ndl
