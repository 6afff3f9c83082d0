---------------------------- MODULE MCAsyncIsr ----------------------------
(* Model-checking wrapper for AsyncIsr.tla (the reference ships no model files).
   AsyncIsr is unbounded as written: LeaderWrite (AsyncIsr.tla:117-119) has no guard and
   the version fields are Nat (AsyncIsr.tla:42-55), so a state CONSTRAINT is required.
   Layout gives every variable a finite type; it is wider by one than Bound because a
   successor that violates the constraint is still generated (and invariant-checked)
   before it is discarded.  pendingVersion holds Nil = -1 initially (AsyncIsr.tla:146),
   which is why AsyncIsr!TypeOk itself cannot serve as the layout. *)
EXTENDS AsyncIsr

CONSTANT MaxVersion

Bound ==
    /\ leaderState.offsets[Leader] <= MaxOffset
    /\ controllerState.version <= MaxVersion

Versions == 0 .. (MaxVersion + 1)
WideOffsets == 0 .. (MaxOffset + 1)
BoundedMessage == [isr : SUBSET Replicas, version : Versions]

Layout ==
    /\ controllerState \in [isr : SUBSET Replicas, version : Versions]
    /\ leaderState \in [isr : SUBSET Replicas,
                        version : Versions,
                        pendingIsr : SUBSET Replicas,
                        pendingVersion : Versions \union {Nil},
                        offsets : [Replicas -> WideOffsets]]
    /\ requests \subseteq BoundedMessage
    /\ updates \subseteq BoundedMessage
=============================================================================
