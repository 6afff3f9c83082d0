--------------------------- MODULE MCKip320Sym ---------------------------
(* Kip320 with TLC symmetry reduction over the replica set: every definition of the Kafka family
   uses replicas only through equality and set membership (KafkaReplication.tla:126-131,158-179),
   so Replicas is a symmetric set of model values.  Permutations is TLC's standard operator. *)
EXTENDS Kip320, TLC
Symm == Permutations(Replicas)
=============================================================================
