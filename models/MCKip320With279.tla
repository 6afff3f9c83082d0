--------------------------- MODULE MCKip320With279 ---------------------------
(* The experiment the reference itself proposes (Kip320.tla:126-133): "Without it [the leader/follower epoch
   check], we violate the strong ISR property ... You can verify this failure by replacing this action with
   `BecomeFollowerTruncateKip279` in the spec below."  NextWith279 is Kip320's Next (Kip320.tla:150-159) with
   FencedBecomeFollowerAndTruncate replaced by Kip279's BecomeFollowerTruncateKip279 (Kip279.tla:47-51). *)
EXTENDS Kip320
NextWith279 ==
    \/ ControllerElectLeader
    \/ ControllerShrinkIsr
    \/ BecomeLeader
    \/ FencedLeaderExpandIsr
    \/ FencedLeaderShrinkIsr
    \/ LeaderWrite
    \/ FencedLeaderIncHighWatermark
    \/ BecomeFollowerTruncateKip279
    \/ FencedFollowerFetch
=============================================================================
