"""`Sequencer`: host-side mirror of the reference's polyphony manager (src/sequencer.rs) — units scheduled with sample-accurate
start and end times and fade envelopes, mixed into one output.

Like `An` and `Net`, a `Sequencer` here is a description: the list of pushed events and edits, replayed onto a backend by
`lower(backend)` (`sequencer_new / sequencer_push / sequencer_edit`). On the GPU every event is one voice of a bank: the voice's
program is the event's graph wrapped in the device node `Event<X>`, which carries the reference's per-block scheduling arithmetic
(f64 time, activation threshold, start / end index, fade phase) itself, so a sequence renders in long launches with no host
involvement per block (fundsp_b200/bank.py `GpuBank.from_sequencer`).
"""
from __future__ import annotations

from .graph import An, _arity


class Fade:  # src/sequencer.rs:35-52
    Power = 0
    Smooth = 1


class ReplayMode:  # src/sequencer.rs:219-229
    All = (0, 0.0)
    None_ = (1, 0.0)

    @staticmethod
    def Loop(t):
        return (2, float(t))


def event(unit: An, start_time, end_time, fade_ease=Fade.Smooth, fade_in_time=0.0, fade_out_time=0.0, loop=0.0):
    """One sequencer event as a voice graph (what `Sequencer.voices()` is made of). `loop`: the period of a `ReplayMode::Loop(loop)` sequencer."""
    _arity(unit.inputs() == 0, "event: the GPU event renders generators (inputs == 0)")
    if loop:
        return An("event_loop", (float(start_time), float(end_time), int(fade_ease), float(fade_in_time), float(fade_out_time), float(loop)), (unit,), 0, unit.outputs())
    return An("event", (float(start_time), float(end_time), int(fade_ease), float(fade_in_time), float(fade_out_time)), (unit,), 0, unit.outputs())


def slot(unit: An):
    """`Slot::new(unit)` as a voice graph: a unit that `GpuBank.slot_set` can replace with a crossfade (src/slot.rs)."""
    return An("slot", (), (unit,), unit.inputs(), unit.outputs())


class Sequencer:
    def __init__(self, inputs, outputs, mode=ReplayMode.None_):  # Sequencer::new (src/sequencer.rs:272-300)
        self.nin, self.nout, self.mode = int(inputs), int(outputs), mode
        self.events = []   # (start, end, fade_ease, fade_in, fade_out, unit, relative)
        self.edits = []    # (event index, end_time, fade_out, relative), applied after the pushes in order
        self.log = []      # pushes and edits in call order: ("push", event index) / ("edit", edit index)

    def inputs(self):
        return self.nin

    def outputs(self):
        return self.nout

    def push(self, start_time, end_time, fade_ease, fade_in_time, fade_out_time, unit: An, _relative=False):  # :319-345
        _arity(unit.inputs() == self.nin and unit.outputs() == self.nout, "sequencer.push: unit arity differs from the sequencer's")
        duration = end_time - start_time
        assert fade_in_time <= duration and fade_out_time <= duration
        self.events.append((float(start_time), float(end_time), int(fade_ease), float(fade_in_time), float(fade_out_time), unit, _relative))
        self.log.append(("push", len(self.events) - 1))
        return len(self.events) - 1     # EventId

    def push_relative(self, start_time, end_time, fade_ease, fade_in_time, fade_out_time, unit: An):  # :376-400 (time is 0 while describing)
        return self.push(start_time, end_time, fade_ease, fade_in_time, fade_out_time, unit, _relative=True)

    def push_duration(self, start_time, duration, fade_ease, fade_in_time, fade_out_time, unit: An):  # :420-437
        return self.push(start_time, start_time + duration, fade_ease, fade_in_time, fade_out_time, unit)

    def edit(self, event_id, end_time, fade_out_time, _relative=False):  # :441-483
        self.edits.append((int(event_id), float(end_time), float(fade_out_time), _relative))
        self.log.append(("edit", len(self.edits) - 1))

    def edit_relative(self, event_id, end_time, fade_out_time):  # :486-528
        self.edit(event_id, end_time, fade_out_time, _relative=True)

    def voices(self):
        """The events as voice graphs for a GPU bank (`GpuBank.from_sequencer`), in push order, with the edits recorded so far applied
        the way the reference applies the edit of a not-yet-active event (its end time and fade-out are replaced, :531-553).
        Generators only. ReplayMode::None or All: a bank reset replays everything, which is ReplayMode::All. ReplayMode::Loop(t): every
        event carries the loop period and wraps on the device (`Event<X>` in nodes.cuh); an edit recorded before the first render changes
        the event's end for the first pass only in the reference (it sets `end_time`, not `original_end_time`, :531-553), which the
        device event does not model: edits of a looping sequencer are refused."""
        _arity(self.nin == 0, "Sequencer.voices: the GPU sequencer renders generators (inputs == 0)")
        loop = self.mode[1] if self.mode[0] == 2 else 0.0
        _arity(not (loop and self.edits), "Sequencer.voices: edits of a ReplayMode::Loop sequencer are not lowered to the GPU")
        ev = [list(e) for e in self.events]
        for k, end, fo, rel in self.edits:
            ev[k][1], ev[k][4] = end, fo     # time is 0 while describing: relative == absolute
        return [event(u, s, e, ease, fi, fo, loop=loop) for s, e, ease, fi, fo, u, _ in ev]

    def node(self):
        """This sequencer as a graph node (`Net::wrap(Box::new(sequencer))` / a boxed AudioUnit inside an expression)."""
        return An("seqnode", (self,), (), self.nin, self.nout)

    def lower(self, backend):
        h = backend.sequencer_new(self.nin, self.nout, self.mode[0], self.mode[1])
        ids = {}
        for kind, k in self.log:
            if kind == "push":
                s, e, ease, fi, fo, unit, rel = self.events[k]
                ids[k] = backend.sequencer_push(h, s, e, ease, fi, fo, unit.lower(backend), rel)
            else:
                ev, end, fo, rel = self.edits[k]
                backend.sequencer_edit(h, ids[ev], end, fo, rel)
        return h


class GpuSequencer:
    """`Sequencer::new(0, outputs, mode)` whose rendering side is a GPU bank (the role of `SequencerBackend`): the reference's calls
    (`push`, `push_relative`, `push_duration`, `edit`, `edit_relative`, `time`, `reset`, `process`) with the reference's meaning.

    Events pushed before the first render become the bank's voices. An event pushed while the sequencer runs takes over the voice
    of a FINISHED event of the same graph class (`fdsp_bank_push_event`); `reserve(unit, count)` adds spare voices of a class up
    front so that note-ons always find one (a spare is an event that ended at time 0)."""

    def __init__(self, outputs, mode=ReplayMode.None_, device=0, sample_rate=44100.0):
        self.nout, self.mode, self.device, self.sr = int(outputs), mode, device, float(sample_rate)
        self.loop = float(mode[1]) if mode[0] == 2 else 0.0     # ReplayMode::Loop(t): events pushed before the first render loop on the device
        self.pending = []      # event expressions until the bank exists
        self.bank = None
        self.voice_of = {}     # EventId -> voice
        self.next_id = 0

    def inputs(self): return 0
    def outputs(self): return self.nout

    def time(self):
        return self.bank.time() if self.bank is not None else 0.0

    def reserve(self, unit: An, count):
        assert self.bank is None, "reserve before the first render"
        for _ in range(int(count)):
            self.pending.append(event(unit, 0.0, 0.0))

    def _push(self, start, end, ease, fade_in, fade_out, unit):
        _arity(unit.inputs() == 0 and unit.outputs() == self.nout, "sequencer.push: unit arity differs from the sequencer's")
        assert fade_in <= end - start and fade_out <= end - start
        _arity(not (self.loop and self.bank is not None), "GpuSequencer: a looping sequencer takes its events before the first render")
        ev = event(unit, start, end, ease, fade_in, fade_out, loop=self.loop)
        eid = self.next_id; self.next_id += 1
        if self.bank is None:
            self.pending.append(ev); self.voice_of[eid] = len(self.pending) - 1
        elif self.mode[0] == 1:
            v = self.bank.push_event(ev)                       # ReplayMode::None drops finished events: their voices are free to take over
            # the finished event that owned this voice is gone for good: its id must not reach the new tenant (the reference ignores
            # edits of past events in this mode, src/sequencer.rs:466-476)
            for old in [k for k, w in self.voice_of.items() if w == v]:
                del self.voice_of[old]
            self.voice_of[eid] = v
        else:
            self.voice_of[eid] = self.bank.add_voice(ev)       # ReplayMode::All replays every event after a reset: nothing may be overwritten
        return eid

    def push(self, start_time, end_time, fade_ease, fade_in_time, fade_out_time, unit: An):
        return self._push(float(start_time), float(end_time), fade_ease, float(fade_in_time), float(fade_out_time), unit)

    def push_relative(self, start_time, end_time, fade_ease, fade_in_time, fade_out_time, unit: An):
        t = self.time()
        return self._push(start_time + t, end_time + t, fade_ease, float(fade_in_time), float(fade_out_time), unit)

    def push_duration(self, start_time, duration, fade_ease, fade_in_time, fade_out_time, unit: An):
        return self.push(start_time, start_time + duration, fade_ease, fade_in_time, fade_out_time, unit)

    def edit(self, event_id, end_time, fade_out_time):
        _arity(not self.loop, "GpuSequencer: edits of a ReplayMode::Loop sequencer are not lowered to the GPU")
        v = self.voice_of.get(event_id)
        if v is None:
            return             # an unknown or past event: a no-op, like the reference's edit
        if self.bank is None:
            a = self.pending[v].args
            self.pending[v] = An("event", (a[0], float(end_time), a[2], a[3], float(fade_out_time)), self.pending[v].kids, 0, self.nout)
        else:
            self.bank.edit_event(v, end_time, fade_out_time)

    def edit_relative(self, event_id, end_time, fade_out_time):
        self.edit(event_id, self.time() + end_time, fade_out_time)

    def _ensure(self):
        if self.bank is None:
            from .bank import GpuBank
            _arity(len(self.pending) > 0, "GpuSequencer: push or reserve at least one event before rendering")
            self.bank = GpuBank(self.pending, device=self.device, per_voice=False, mix=True, sample_rate=self.sr)
        return self.bank

    def reset(self):
        """ReplayMode::All: every event replays. ReplayMode::None: the sequencer is emptied (its voices stay as spares)."""
        b = self._ensure()
        b.reset()
        if self.mode[0] == 1:
            for v in range(b.voices()):
                b.edit_event(v, 0.0, 0.0)
            b.reset()          # the edited (ended) events are now the construction-time state as well
            self.voice_of.clear()

    def process(self, size):
        return self._ensure().process(size)

    def render(self, n):
        return self._ensure().render_samples(int(n))[1]
